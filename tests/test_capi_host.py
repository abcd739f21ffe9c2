"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/jrlqp_b200.h declares, and fails loudly (no CPU fallback) when no B200 is present."""
import ctypes as C
import os
import re

import pytest
import torch

import jrl_qp_b200  # noqa: F401
from jrl_qp_b200 import solver as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    for fn in os.listdir(os.path.join(ROOT, "include")):
        if fn.endswith(".h"):
            text = open(os.path.join(ROOT, "include", fn)).read()
            text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
            names |= set(re.findall(r"\b(jrlqp_[a-z0-9_]+)\s*\(", text))
    return names


def test_library_exports_every_declared_symbol():
    lib = S.load_library()
    declared = _declared_symbols()
    assert declared, "no declarations found in include/*.h"
    assert set(S.EXPORTED_SYMBOLS) <= declared
    for name in sorted(declared):
        assert hasattr(lib, name), f"libjrlqp_b200.so does not export {name}"
    assert lib.jrlqp_version() == S.ABI_VERSION


def test_struct_mirrors_match_header_sizes():
    # jrlqp_problem: 1 + 7*2 + 2 (ld) + 2 (as_in) fields; sizes follow the x86-64 C layout
    assert C.sizeof(S._Options) == 24
    assert C.sizeof(S._Problem) == 8 * 19
    assert C.sizeof(S._Result) == 8 * 9
    assert C.sizeof(S.KernelInfo) == 4 * 8
    assert C.sizeof(S._Sequence) == 72  # 2 x int32, 6 x int64, 2 pointers
    assert C.sizeof(S._KktArgs) == 88  # 3 x int32 (+4 pad), 3 doubles, 6 pointers


def test_default_options_match_reference():
    lib = S.load_library()
    o = S._Options()
    lib.jrlqp_default_options(C.byref(o))
    # include/jrl-qp/SolverOptions.h:16-19
    assert (o.max_iter, o.big_bnd, o.warm_start, o.log_flags) == (500, 1e100, 0, 0)
    py = S.SolverOptions()
    assert (py.maxIter(), py.bigBnd(), py.warmStart(), py.logFlags()) == (500, 1e100, False, 0)
    assert py.maxIter(10).bigBnd(1e30).warmStart(True) is py and py.maxIter() == 10


def test_invalid_arguments_are_rejected():
    lib = S.load_library()
    h = C.c_void_p()
    assert lib.jrlqp_create(C.byref(h), 0, 1, 0, 1, 0) == -2  # JRLQP_ERR_ARG
    assert lib.jrlqp_create(C.byref(h), 1025, 1, 0, 1, 0) == -2
    assert lib.jrlqp_create(None, 5, 1, 0, 1, 0) == -2


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_without_gpu():
    with pytest.raises(S.JrlQpError):
        S.BatchedGoldfarbIdnaniSolver(5, 3, True, 4)
    with pytest.raises(S.JrlQpError):
        S.GoldfarbIdnaniSolver(5, 3, True)
    assert S.measure_fp64_tflops() < 0


def test_enums_match_reference_values():
    # include/jrl-qp/enums.h:14-37
    assert [s.name for s in S.ActivationStatus] == ["INACTIVE", "LOWER", "UPPER", "EQUALITY", "LOWER_BOUND",
                                                    "UPPER_BOUND", "FIXED"]
    assert [s.name for s in S.TerminationStatus] == ["SUCCESS", "INCONSISTENT_INPUT", "NON_POS_HESSIAN", "INFEASIBLE",
                                                     "MAX_ITER_REACHED", "LINEAR_DEPENDENCY_DETECTED",
                                                     "OVERCONSTRAINED_PROBLEM", "UNKNOWN"]
