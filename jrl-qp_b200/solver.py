"""Python mirror of the reference's solver interface over the C-ABI (include/jrlqp_b200.h).

`GoldfarbIdnaniSolver` keeps the reference's method names and argument meaning
(include/jrl-qp/GoldfarbIdnaniSolver.h:15-33, include/jrl-qp/DualSolver.h:26-60):
solve(G, a, C, bl, bu, xl, xu) -> TerminationStatus, solution(), multipliers(), objectiveValue(),
iterations(), activeSet(), resetActiveSet(), options(), resize(). `BatchedGoldfarbIdnaniSolver` is the
same call over a batch. Both go through ctypes to libjrlqp_b200.so — there is no CPU fallback: if the
CUDA library is missing or no B200 is visible, construction raises.
"""
import ctypes as C
import enum
import os

import numpy as np

from . import build as _build


class ActivationStatus(enum.IntEnum):  # include/jrl-qp/enums.h:14-23
    INACTIVE = 0
    LOWER = 1
    UPPER = 2
    EQUALITY = 3
    LOWER_BOUND = 4
    UPPER_BOUND = 5
    FIXED = 6


class TerminationStatus(enum.IntEnum):  # include/jrl-qp/enums.h:26-37
    SUCCESS = 0
    INCONSISTENT_INPUT = 1
    NON_POS_HESSIAN = 2
    INFEASIBLE = 3
    MAX_ITER_REACHED = 4
    LINEAR_DEPENDENCY_DETECTED = 5
    OVERCONSTRAINED_PROBLEM = 6
    UNKNOWN = 7


class SolverOptions:
    """include/jrl-qp/SolverOptions.h:14-88 (chained setters; log stream not carried)."""

    def __init__(self):
        self.maxIter_ = 500
        self.bigBnd_ = 1e100
        self.warmStart_ = False
        self.logFlags_ = 0

    def maxIter(self, v=None):
        if v is None:
            return self.maxIter_
        self.maxIter_ = int(v)
        return self

    def bigBnd(self, v=None):
        if v is None:
            return self.bigBnd_
        self.bigBnd_ = float(v)
        return self

    def warmStart(self, v=None):
        if v is None:
            return self.warmStart_
        self.warmStart_ = bool(v)
        return self

    def logFlags(self, v=None):
        if v is None:
            return self.logFlags_
        self.logFlags_ = int(v)
        return self


class _Options(C.Structure):
    _fields_ = [("max_iter", C.c_int32), ("big_bnd", C.c_double), ("warm_start", C.c_int32), ("log_flags", C.c_uint32)]


class _Problem(C.Structure):
    _fields_ = [("batch", C.c_int64),
                ("G", C.c_void_p), ("G_stride", C.c_int64), ("ldg", C.c_int32),
                ("a", C.c_void_p), ("a_stride", C.c_int64),
                ("C", C.c_void_p), ("C_stride", C.c_int64), ("ldc", C.c_int32),
                ("bl", C.c_void_p), ("bl_stride", C.c_int64),
                ("bu", C.c_void_p), ("bu_stride", C.c_int64),
                ("xl", C.c_void_p), ("xl_stride", C.c_int64),
                ("xu", C.c_void_p), ("xu_stride", C.c_int64),
                ("as_in", C.c_void_p), ("as_stride", C.c_int64)]


class _Result(C.Structure):
    _fields_ = [("x", C.c_void_p), ("u", C.c_void_p), ("f", C.c_void_p), ("iterations", C.c_void_p),
                ("status", C.c_void_p), ("active_set", C.c_void_p), ("active_list", C.c_void_p),
                ("n_active", C.c_void_p), ("L", C.c_void_p)]


class _Sequence(C.Structure):
    _fields_ = [("steps", C.c_int32), ("warm", C.c_int32), ("a_step_stride", C.c_int64), ("x_step_stride", C.c_int64),
                ("u_step_stride", C.c_int64), ("f_step_stride", C.c_int64), ("iterations_step_stride", C.c_int64),
                ("status_step_stride", C.c_int64), ("iterations_total", C.c_void_p), ("status_worst", C.c_void_p)]


class _KktArgs(C.Structure):
    _fields_ = [("n", C.c_int32), ("mc", C.c_int32), ("use_bounds", C.c_int32), ("tau_p", C.c_double), ("tau_d", C.c_double),
                ("prec", C.c_double), ("x", C.c_void_p), ("u", C.c_void_p), ("x_ref", C.c_void_p), ("flags", C.c_void_p),
                ("resid", C.c_void_p), ("n_fail", C.c_void_p)]


class KernelInfo(C.Structure):
    _fields_ = [("threads_per_qp", C.c_int32), ("rows_per_thread", C.c_int32), ("smem_bytes_per_qp", C.c_int32),
                ("qps_per_sm", C.c_int32), ("grid", C.c_int32), ("num_sms", C.c_int32), ("stage_c", C.c_int32),
                ("regs_per_thread", C.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


EXPORTED_SYMBOLS = [
    "jrlqp_version", "jrlqp_default_options", "jrlqp_create", "jrlqp_destroy", "jrlqp_set_options",
    "jrlqp_get_options", "jrlqp_solve_batch_device", "jrlqp_solve_batch_host", "jrlqp_get_kernel_info",
    "jrlqp_set_stage_c", "jrlqp_launch_count", "jrlqp_last_error", "jrlqp_measure_fp64_tflops",
    "jrlqp_structured_create", "jrlqp_structured_destroy", "jrlqp_structured_last_error",
    "jrlqp_structured_llt_device", "jrlqp_structured_llt_host", "jrlqp_structured_solve_device",
    "jrlqp_structured_solve_host", "jrlqp_structured_get_info", "jrlqp_structured_set_kernel", "jrlqp_selftest_arith",
    "jrlqp_solve_batch_warm_device", "jrlqp_solve_batch_warm_host", "jrlqp_set_kernel_path", "jrlqp_set_scan_transposed", "jrlqp_host_g_bytes",
    "jrlqp_solve_sequence_device", "jrlqp_solve_sequence_host",
    "jrlqp_kkt_default_args", "jrlqp_kkt_check_device", "jrlqp_kkt_check_host",
    "jrlqp_blockgi_create", "jrlqp_blockgi_destroy", "jrlqp_blockgi_last_error", "jrlqp_blockgi_set_options",
    "jrlqp_blockgi_get_options", "jrlqp_blockgi_solve_device", "jrlqp_blockgi_solve_host", "jrlqp_blockgi_get_info",
    "jrlqp_multi_create", "jrlqp_multi_destroy", "jrlqp_multi_set_options", "jrlqp_multi_device_count", "jrlqp_multi_device",
    "jrlqp_multi_solver", "jrlqp_multi_shard", "jrlqp_multi_solve_batch_host", "jrlqp_multi_solve_batch_warm_host",
    "jrlqp_blockgi_test_sequence", "jrlqp_probe_dmma", "jrlqp_multi_last_error", "jrlqp_multi_set_balancing", "jrlqp_multi_get_weights", "jrlqp_measure_host_link",
]

ABI_VERSION = 200  # JRLQP_B200_VERSION of include/jrlqp_b200.h the ctypes structures below were written against

_check_abi = True  # scripts/ab_variants.py loads older variant builds (same structures) on purpose
_lib = None


def library_path():
    if os.environ.get("JRLQP_B200_LIB"):  # development aid: a variant build of the same sources (scripts/build_variant.py)
        return os.environ["JRLQP_B200_LIB"]
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "libjrlqp_b200.so")


def load_library():
    """Load libjrlqp_b200.so (building it in-tree if it is missing). Raises if that fails."""
    global _lib
    if _lib is None:
        path = library_path()
        if not os.environ.get("JRLQP_B200_LIB"):
            # incremental (per-object staleness against the sources and the headers they include): a stale binary whose
            # structures no longer match the ctypes mirrors below is never loaded silently. Without nvcc (a box that
            # only received the built library) the existing binary is used and checked by its version alone.
            try:
                _build.build_cuda()
            except (OSError, RuntimeError):
                if not os.path.exists(path):
                    raise
        lib = C.CDLL(path)
        if _check_abi and not os.environ.get("JRLQP_B200_LIB") and lib.jrlqp_version() != ABI_VERSION:  # (a variant named by the environment is a development aid)
            raise ImportError(f"{path}: ABI version {lib.jrlqp_version()}, this package expects {ABI_VERSION} — rebuild (python __graft_entry__.py)")
        lib.jrlqp_launch_count.restype = C.c_int64
        lib.jrlqp_last_error.restype = C.c_char_p
        lib.jrlqp_last_error.argtypes = [C.c_void_p]
        lib.jrlqp_measure_fp64_tflops.restype = C.c_double
        lib.jrlqp_create.argtypes = [C.POINTER(C.c_void_p), C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int32]
        lib.jrlqp_destroy.argtypes = [C.c_void_p]
        lib.jrlqp_set_options.argtypes = [C.c_void_p, C.POINTER(_Options)]
        lib.jrlqp_get_kernel_info.argtypes = [C.c_void_p, C.POINTER(KernelInfo)]
        lib.jrlqp_set_stage_c.argtypes = [C.c_void_p, C.c_int32]
        lib.jrlqp_set_kernel_path.argtypes = [C.c_void_p, C.c_int32]
        lib.jrlqp_set_scan_transposed.argtypes = [C.c_void_p, C.c_int32]
        lib.jrlqp_host_g_bytes.argtypes = [C.c_void_p, C.c_int32]
        lib.jrlqp_host_g_bytes.restype = C.c_int64
        lib.jrlqp_solve_batch_device.argtypes = [C.c_void_p, C.POINTER(_Problem), C.POINTER(_Result), C.c_void_p]
        lib.jrlqp_solve_batch_host.argtypes = [C.c_void_p, C.POINTER(_Problem), C.POINTER(_Result)]
        lib.jrlqp_solve_batch_warm_device.argtypes = [C.c_void_p, C.POINTER(_Problem), C.POINTER(_Result), C.c_void_p]
        lib.jrlqp_solve_batch_warm_host.argtypes = [C.c_void_p, C.POINTER(_Problem), C.POINTER(_Result)]
        lib.jrlqp_solve_sequence_device.argtypes = [C.c_void_p, C.POINTER(_Problem), C.POINTER(_Sequence), C.POINTER(_Result), C.c_void_p]
        lib.jrlqp_solve_sequence_host.argtypes = [C.c_void_p, C.POINTER(_Problem), C.POINTER(_Sequence), C.POINTER(_Result)]
        lib.jrlqp_kkt_default_args.argtypes = [C.POINTER(_KktArgs)]
        lib.jrlqp_kkt_default_args.restype = None
        lib.jrlqp_kkt_check_device.argtypes = [C.POINTER(_Problem), C.POINTER(_KktArgs), C.c_int32, C.c_void_p]
        lib.jrlqp_kkt_check_host.argtypes = [C.POINTER(_Problem), C.POINTER(_KktArgs), C.c_int32]
        lib.jrlqp_structured_create.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_int64, C.c_int32]
        lib.jrlqp_structured_destroy.argtypes = [C.c_void_p]
        lib.jrlqp_structured_last_error.restype = C.c_char_p
        lib.jrlqp_structured_last_error.argtypes = [C.c_void_p]
        lib.jrlqp_structured_get_info.argtypes = [C.c_void_p, C.c_void_p]
        lib.jrlqp_structured_llt_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
        lib.jrlqp_structured_llt_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]
        lib.jrlqp_structured_solve_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int32,
                                                      C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
        lib.jrlqp_structured_solve_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int32,
                                                    C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32]
        lib.jrlqp_multi_create.argtypes = [C.POINTER(C.c_void_p), C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_int32]
        lib.jrlqp_multi_destroy.argtypes = [C.c_void_p]
        lib.jrlqp_multi_set_options.argtypes = [C.c_void_p, C.POINTER(_Options)]
        lib.jrlqp_multi_device_count.argtypes = [C.c_void_p]
        lib.jrlqp_multi_device.argtypes = [C.c_void_p, C.c_int32]
        lib.jrlqp_multi_solver.argtypes = [C.c_void_p, C.c_int32]
        lib.jrlqp_multi_solver.restype = C.c_void_p
        lib.jrlqp_multi_shard.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        lib.jrlqp_multi_solve_batch_host.argtypes = [C.c_void_p, C.POINTER(_Problem), C.POINTER(_Result)]
        lib.jrlqp_multi_solve_batch_warm_host.argtypes = [C.c_void_p, C.POINTER(_Problem), C.POINTER(_Result)]
        lib.jrlqp_multi_set_balancing.argtypes = [C.c_void_p, C.c_int32]
        lib.jrlqp_multi_get_weights.argtypes = [C.c_void_p, C.c_void_p]
        lib.jrlqp_multi_last_error.argtypes = [C.c_void_p]
        lib.jrlqp_multi_last_error.restype = C.c_char_p
        lib.jrlqp_measure_host_link.argtypes = [C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_void_p]
        lib.jrlqp_measure_host_link.restype = C.c_double
        _lib = lib
    return _lib


def measure_host_link(n_devices, nbytes=1 << 30, reps=4, direction=0, devices=None):
    """Aggregate and per-device GB/s of concurrent pinned-host <-> device copies (jrlqp_measure_host_link):
    direction 0 host -> device, 1 device -> host, 2 both at once."""
    lib = load_library()
    per = (C.c_double * n_devices)()
    devs = None if devices is None else (C.c_int32 * n_devices)(*devices)
    agg = lib.jrlqp_measure_host_link(devs, n_devices, nbytes, reps, direction, per)
    if agg < 0:
        raise JrlQpError("jrlqp_measure_host_link failed")
    return float(agg), [float(v) for v in per]


def probe_dmma(device=0, reps=64):
    """jrlqp_probe_dmma: dict of the six figures of the FP64 tensor-core experiment."""
    out = (C.c_double * 6)()
    rc = load_library().jrlqp_probe_dmma(C.c_int32(device), C.c_int32(reps), out)
    if rc != 0:
        raise JrlQpError(f"jrlqp_probe_dmma failed ({rc})")
    keys = ("fp64_pipe_gflops", "dmma_gflops", "dmma_equals_sequential_fma_chain", "dmma_equals_pairwise_order",
            "dmma_dot4_interleaving_equals_dot4", "pipe_and_dmma_products_agree")
    return dict(zip(keys, [float(v) for v in out]))


def selftest_arith(samples=1 << 24, seed=1, exponent_span=30, rcp_ulps=3, device=0):
    """counts of jrlqp_selftest_arith: (proven, proven_but_wrong, unproven, sqrt_mismatch, rsqrt_far)."""
    lib = load_library()
    out = (C.c_uint64 * 5)()
    rc = lib.jrlqp_selftest_arith(C.c_int32(device), C.c_int64(samples), C.c_uint64(seed), C.c_int32(exponent_span),
                                  C.c_int32(rcp_ulps), out)
    if rc != 0:
        raise RuntimeError(f"jrlqp_selftest_arith failed ({rc})")
    return tuple(int(v) for v in out)


def launch_count():
    return int(load_library().jrlqp_launch_count())


def measure_fp64_tflops(device=0, repeats=5):
    return float(load_library().jrlqp_measure_fp64_tflops(C.c_int32(device), C.c_int32(repeats)))


class JrlQpError(RuntimeError):
    pass


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()  # torch tensor (host pinned or device)


class BatchedGoldfarbIdnaniSolver:
    """Batched GoldfarbIdnaniSolver. Arrays follow the C-ABI layout:
    G [B,n,n] column-major blocks, a [B,n], C [B,mc,n] (row i = normal of constraint i, which IS the
    reference's n x mc column-major matrix), bl/bu [B,mc], xl/xu [B,n] or None. An array with one
    dimension less is shared by the whole batch (stride 0)."""

    def __init__(self, nbVar, nbCstr, useBounds, batch_capacity=1, device=0):
        self._lib = load_library()
        self.n, self.mc, self.nb = int(nbVar), int(nbCstr), (int(nbVar) if useBounds else 0)
        self.m = self.mc + self.nb
        self.capacity = int(batch_capacity)
        self.device = device
        self._h = C.c_void_p()
        rc = self._lib.jrlqp_create(C.byref(self._h), self.n, self.mc, int(bool(useBounds)), self.capacity, device)
        if rc != 0:
            msg = self._lib.jrlqp_last_error(self._h).decode() if self._h else "allocation failed"
            if self._h:
                self._lib.jrlqp_destroy(self._h)
                self._h = None
            raise JrlQpError(f"jrlqp_create failed ({rc}): {msg}")
        self._options = SolverOptions()
        self.last = None

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._lib.jrlqp_destroy(h)
            self._h = None

    # DualSolver::options
    def options(self, opt=None):
        if opt is None:
            return self._options
        self._options = opt
        o = _Options(opt.maxIter_, opt.bigBnd_, int(opt.warmStart_), opt.logFlags_)
        rc = self._lib.jrlqp_set_options(self._h, C.byref(o))
        if rc != 0:
            raise JrlQpError("jrlqp_set_options failed")
        return self

    def kernel_info(self):
        info = KernelInfo()
        self._lib.jrlqp_get_kernel_info(self._h, C.byref(info))
        return info.as_dict()

    def set_stage_c(self, mode):
        rc = self._lib.jrlqp_set_stage_c(self._h, int(mode))
        if rc != 0:
            raise JrlQpError(f"jrlqp_set_stage_c failed: {self._lib.jrlqp_last_error(self._h).decode()}")

    def host_g_bytes(self, pinned=True):
        """Bytes of one instance's G that cross the host link in solve() (jrlqp_host_g_bytes)."""
        return int(self._lib.jrlqp_host_g_bytes(self._h, int(bool(pinned))))

    def set_scan_transposed(self, on):
        """Constraint scan of the non-staged shared-memory kernels: transposed copy of C (default) or C in place."""
        if self._lib.jrlqp_set_scan_transposed(self._h, -1 if on is None else int(bool(on))) != 0:
            raise JrlQpError("jrlqp_set_scan_transposed failed")
        return self

    def set_kernel_path(self, mode):
        """0 automatic, 1 shared-memory kernel (n <= 128), 2 global-workspace kernel (any n <= 1024)."""
        rc = self._lib.jrlqp_set_kernel_path(self._h, int(mode))
        if rc != 0:
            raise JrlQpError(f"jrlqp_set_kernel_path failed: {self._lib.jrlqp_last_error(self._h).decode()}")
        return self

    def _problem(self, B, G, a, Cm, bl, bu, xl, xu, shared, ldg=None, ldc=None):
        n, mc = self.n, self.mc
        st = lambda name, per: 0 if name in shared else per
        pb = _Problem()
        pb.batch = B
        pb.G, pb.G_stride, pb.ldg = _ptr(G), st("G", n * n), ldg or n
        pb.a, pb.a_stride = _ptr(a), st("a", n)
        pb.C, pb.C_stride, pb.ldc = _ptr(Cm), st("C", mc * n), ldc or n
        pb.bl, pb.bl_stride = _ptr(bl), st("bl", mc)
        pb.bu, pb.bu_stride = _ptr(bu), st("bu", mc)
        pb.xl, pb.xl_stride = _ptr(xl), st("xl", n)
        pb.xu, pb.xu_stride = _ptr(xu), st("xu", n)
        pb.as_in, pb.as_stride = None, 0
        return pb

    def solve(self, G, a, Cm, bl, bu, xl=None, xu=None, want_L=False, want_active_list=True, experimental=False, as_in=None):
        """HOST arrays (numpy). Returns the worst TerminationStatus; results in self.last (dict).
        experimental=True runs the warm-start capable solver (experimental::GoldfarbIdnaniSolver::solve):
        as_in [B, mc+nb] int8 (or [mc+nb], shared) is the guessed active set, honoured when
        options().warmStart() is set."""
        f64 = lambda v: None if v is None else np.ascontiguousarray(v, dtype=np.float64)
        G, a, Cm, bl, bu, xl, xu = map(f64, (G, a, Cm, bl, bu, xl, xu))
        if self.nb == 0:
            xl = xu = None
        n, mc, m = self.n, self.mc, self.m
        full = {"G": 3, "a": 2, "C": 3, "bl": 2, "bu": 2, "xl": 2, "xu": 2}
        arrs = {"G": G, "a": a, "C": Cm, "bl": bl, "bu": bu, "xl": xl, "xu": xu}
        shared = {k for k, v in arrs.items() if v is not None and v.ndim < full[k]}
        Bs = [v.shape[0] for k, v in arrs.items() if v is not None and k not in shared]
        B = max(Bs) if Bs else 1
        if mc == 0:
            Cm = bl = bu = None
        # sizes are checked here, as the reference's asserts do (src/GoldfarbIdnaniSolver.cpp:33-39): a mis-shaped array
        # would otherwise be read out of bounds by the copies
        want = {"G": (n, n), "a": (n,), "C": (mc, n), "bl": (mc,), "bu": (mc,), "xl": (n,), "xu": (n,)}
        for k, v in (("G", G), ("a", a), ("C", Cm), ("bl", bl), ("bu", bu), ("xl", xl), ("xu", xu)):
            if v is None:
                if k in ("G", "a") or (k in ("C", "bl", "bu") and mc) or (k in ("xl", "xu") and self.nb):
                    raise JrlQpError(f"INCONSISTENT_INPUT: {k} is missing")
                continue
            if tuple(v.shape[-len(want[k]):]) != want[k] or (k not in shared and v.shape[0] != B) or v.ndim not in (full[k], full[k] - 1):
                raise JrlQpError(f"INCONSISTENT_INPUT: {k} has shape {v.shape}, expected {('B',) + want[k]} (or {want[k]} shared) with B = {B}")
        if B > self.capacity:
            raise JrlQpError(f"batch {B} exceeds the capacity {self.capacity} given at construction")
        x = np.empty((B, n))
        u = np.empty((B, m))
        f = np.empty(B)
        it = np.empty(B, dtype=np.int32)
        status = np.empty(B, dtype=np.int32)
        act = np.empty((B, m), dtype=np.int8)
        alist = np.empty((B, n), dtype=np.int32) if want_active_list else None
        nact = np.empty(B, dtype=np.int32)
        L = np.zeros((B, n, n)) if want_L else None
        pb = self._problem(B, G, a, Cm, bl, bu, xl, xu, shared)
        res = _Result(_ptr(x), _ptr(u), _ptr(f), _ptr(it), _ptr(status), _ptr(act), _ptr(alist), _ptr(nact), _ptr(L))
        if experimental:
            if as_in is not None:
                as_in = np.ascontiguousarray(as_in, dtype=np.int8)
                if as_in.shape[-1] != m:
                    raise JrlQpError("as_in must have nbCstr + nbBnd entries per instance")
                pb.as_in, pb.as_stride = _ptr(as_in), (m if as_in.ndim == 2 else 0)
        rc = self._host_call(pb, res, experimental)
        if rc < 0:
            raise JrlQpError(f"jrlqp_solve_batch_host failed ({rc}): {self._last_error()}")
        self.last = dict(x=x, u=u, f=f, iterations=it, status=status, active_set=act, active_list=alist,
                         n_active=nact, worst=rc)
        if want_L:
            self.last["L"] = L
        return TerminationStatus(rc)

    def _host_call(self, pb, res, experimental):
        fn = self._lib.jrlqp_solve_batch_warm_host if experimental else self._lib.jrlqp_solve_batch_host
        return fn(self._h, C.byref(pb), C.byref(res))

    def _last_error(self):
        return self._lib.jrlqp_last_error(self._h).decode()

    def solve_device(self, B, G, a, Cm, bl, bu, xl, xu, x, u=None, f=None, iterations=None, status=None,
                     active_set=None, active_list=None, n_active=None, L=None, stream=0, shared=(), ldg=None, ldc=None,
                     strides=None, experimental=False, as_in=None, as_shared=False):
        """DEVICE pointers (ints or torch CUDA tensors); asynchronous on `stream` (cudaStream_t as int).
        `shared` names arrays with stride 0; `strides` overrides element strides, e.g. {"G": n * ldg}.
        experimental=True: the warm-start capable solver, as_in = device int8 [B, mc+nb] guess (or shared)."""
        pb = self._problem(B, G, a, Cm, bl, bu, xl, xu, set(shared), ldg, ldc)
        for k, v in (strides or {}).items():
            setattr(pb, k + "_stride", int(v))
        res = _Result(_ptr(x), _ptr(u), _ptr(f), _ptr(iterations), _ptr(status), _ptr(active_set),
                      _ptr(active_list), _ptr(n_active), _ptr(L))
        if experimental:
            if as_in is not None:
                pb.as_in, pb.as_stride = _ptr(as_in), (0 if as_shared else self.m)
            rc = self._lib.jrlqp_solve_batch_warm_device(self._h, C.byref(pb), C.byref(res), C.c_void_p(stream))
        else:
            rc = self._lib.jrlqp_solve_batch_device(self._h, C.byref(pb), C.byref(res), C.c_void_p(stream))
        if rc != 0:
            raise JrlQpError(f"jrlqp_solve_batch_device failed ({rc}): {self._lib.jrlqp_last_error(self._h).decode()}")


    def solve_sequence(self, G, a_seq, Cm, bl, bu, xl=None, xu=None, warm=True, as_in=None, keep_steps=False):
        """The reference's warm-start benchmark loop (benchmarks/SolversWarmStart.cpp:234-276) for a whole batch:
        a_seq [T, B, n] are the linear terms of T consecutive steps, G / C / bounds as in solve(). warm=True: the
        experimental solver, every step warm-started from the active set of the previous one (the first from
        as_in, if given); warm=False: T cold solves of the stable solver. HOST arrays. Results in self.last:
        the outputs of the last step (x, u, f, iterations, status [T, B, ...] when keep_steps), plus
        iterations_total and status_worst [B]."""
        f64 = lambda v: None if v is None else np.ascontiguousarray(v, dtype=np.float64)
        G, a_seq, Cm, bl, bu, xl, xu = map(f64, (G, a_seq, Cm, bl, bu, xl, xu))
        if self.nb == 0:
            xl = xu = None
        n, mc, m = self.n, self.mc, self.m
        T, B = a_seq.shape[0], a_seq.shape[1]
        full = {"G": 3, "C": 3, "bl": 2, "bu": 2, "xl": 2, "xu": 2}
        arrs = {"G": G, "C": Cm, "bl": bl, "bu": bu, "xl": xl, "xu": xu}
        shared = {k for k, v in arrs.items() if v is not None and v.ndim < full[k]}
        if mc == 0:
            Cm = bl = bu = None
        lead = (T,) if keep_steps else ()
        x = np.empty(lead + (B, n))
        u = np.empty(lead + (B, m))
        f = np.empty(lead + (B,))
        it = np.empty(lead + (B,), dtype=np.int32)
        status = np.empty(lead + (B,), dtype=np.int32)
        act = np.empty((B, m), dtype=np.int8)
        alist = np.empty((B, n), dtype=np.int32)
        nact = np.empty(B, dtype=np.int32)
        it_total = np.empty(B, dtype=np.int32)
        worst = np.empty(B, dtype=np.int32)
        pb = self._problem(B, G, a_seq, Cm, bl, bu, xl, xu, shared)
        if as_in is not None:
            as_in = np.ascontiguousarray(as_in, dtype=np.int8)
            pb.as_in, pb.as_stride = _ptr(as_in), (m if as_in.ndim == 2 else 0)
        k = 1 if keep_steps else 0
        seq = _Sequence(T, int(bool(warm)), B * n, k * B * n, k * B * m, k * B, k * B, k * B, _ptr(it_total), _ptr(worst))
        res = _Result(_ptr(x), _ptr(u), _ptr(f), _ptr(it), _ptr(status), _ptr(act), _ptr(alist), _ptr(nact), None)
        rc = self._lib.jrlqp_solve_sequence_host(self._h, C.byref(pb), C.byref(seq), C.byref(res))
        if rc < 0:
            raise JrlQpError(f"jrlqp_solve_sequence_host failed ({rc}): {self._lib.jrlqp_last_error(self._h).decode()}")
        self.last = dict(x=x, u=u, f=f, iterations=it, status=status, active_set=act, active_list=alist, n_active=nact,
                         iterations_total=it_total, status_worst=worst, worst=rc)
        return TerminationStatus(rc)

    def solve_sequence_device(self, B, steps, G, a_seq, Cm, bl, bu, xl, xu, x, active_set, u=None, f=None, iterations=None,
                              status=None, iterations_total=None, status_worst=None, warm=True, as_in=None, stream=0,
                              shared=(), a_step_stride=None, step_strides=None):
        """DEVICE pointers; asynchronous on `stream`. a_seq: [steps, B, n]; outputs of the last step unless
        step_strides = {"x": ..., "u": ..., "f": ..., "iterations": ..., "status": ...} (elements) are given."""
        pb = self._problem(B, G, a_seq, Cm, bl, bu, xl, xu, set(shared))
        if as_in is not None:
            pb.as_in, pb.as_stride = _ptr(as_in), self.m
        ss = step_strides or {}
        seq = _Sequence(int(steps), int(bool(warm)), int(a_step_stride if a_step_stride is not None else B * self.n),
                        int(ss.get("x", 0)), int(ss.get("u", 0)), int(ss.get("f", 0)), int(ss.get("iterations", 0)),
                        int(ss.get("status", 0)), _ptr(iterations_total), _ptr(status_worst))
        res = _Result(_ptr(x), _ptr(u), _ptr(f), _ptr(iterations), _ptr(status), _ptr(active_set), None, None, None)
        rc = self._lib.jrlqp_solve_sequence_device(self._h, C.byref(pb), C.byref(seq), C.byref(res), C.c_void_p(stream))
        if rc != 0:
            raise JrlQpError(f"jrlqp_solve_sequence_device failed ({rc}): {self._lib.jrlqp_last_error(self._h).decode()}")

    def _kkt_args(self, x, u, x_ref, flags, resid, n_fail, tau_p, tau_d, prec):
        k = _KktArgs()
        self._lib.jrlqp_kkt_default_args(C.byref(k))
        k.n, k.mc, k.use_bounds = self.n, self.mc, int(self.nb > 0)
        if tau_p is not None:
            k.tau_p = tau_p
        if tau_d is not None:
            k.tau_d = tau_d
        if prec is not None:
            k.prec = prec
        k.x, k.u, k.x_ref, k.flags, k.resid, k.n_fail = _ptr(x), _ptr(u), _ptr(x_ref), _ptr(flags), _ptr(resid), _ptr(n_fail)
        return k

    def test_kkt(self, x, u, G, a, Cm, bl, bu, xl=None, xu=None, x_ref=None, tau_p=None, tau_d=None, prec=None):
        """jrl::qp::test::testKKT (src/test/kkt.cpp:87-195) for a batch, on the GPU. HOST arrays laid out as in
        solve(). Returns (flags [B] int32: bit 0 stationarity, bit 1 feasibility, bit 2 x ~ x_ref; resid [B, 4]:
        |dL|_inf, tau_u, tau_x, |x - x_ref|^2; number of failing instances)."""
        f64 = lambda v: None if v is None else np.ascontiguousarray(v, dtype=np.float64)
        G, a, Cm, bl, bu, xl, xu, x, u, x_ref = map(f64, (G, a, Cm, bl, bu, xl, xu, x, u, x_ref))
        if self.nb == 0:
            xl = xu = None
        if self.mc == 0:
            Cm = bl = bu = None
        full = {"G": 3, "a": 2, "C": 3, "bl": 2, "bu": 2, "xl": 2, "xu": 2}
        arrs = {"G": G, "a": a, "C": Cm, "bl": bl, "bu": bu, "xl": xl, "xu": xu}
        shared = {k for k, v in arrs.items() if v is not None and v.ndim < full[k]}
        B = x.shape[0]
        flags = np.empty(B, dtype=np.int32)
        resid = np.empty((B, 4))
        n_fail = np.zeros(1, dtype=np.int64)
        pb = self._problem(B, G, a, Cm, bl, bu, xl, xu, shared)
        k = self._kkt_args(x, u, x_ref, flags, resid, n_fail, tau_p, tau_d, prec)
        rc = self._lib.jrlqp_kkt_check_host(C.byref(pb), C.byref(k), self.device)
        if rc < 0:
            raise JrlQpError(f"jrlqp_kkt_check_host failed ({rc})")
        return flags, resid, int(n_fail[0])

    def test_kkt_device(self, B, x, u, G, a, Cm, bl, bu, xl, xu, flags, resid=None, n_fail=None, x_ref=None, shared=(),
                        tau_p=None, tau_d=None, prec=None, stream=0):
        """Same check on DEVICE pointers, asynchronous on `stream` (n_fail: zeroed device int64 counter)."""
        pb = self._problem(B, G, a, Cm, bl, bu, xl, xu, set(shared))
        k = self._kkt_args(x, u, x_ref, flags, resid, n_fail, tau_p, tau_d, prec)
        rc = self._lib.jrlqp_kkt_check_device(C.byref(pb), C.byref(k), self.device, C.c_void_p(stream))
        if rc != 0:
            raise JrlQpError(f"jrlqp_kkt_check_device failed ({rc})")


class MultiGpuGoldfarbIdnaniSolver(BatchedGoldfarbIdnaniSolver):
    """One host batch over every GPU of the box (jrlqp_multi_*, include/jrlqp_b200.h): contiguous shards, one solver
    and one host thread per device, results written in place — no collective. solve() takes the same
    HOST arrays as BatchedGoldfarbIdnaniSolver.solve and returns the same outputs, bit for bit."""

    def __init__(self, nbVar, nbCstr, useBounds, batch_capacity=1, devices=None, n_devices=0):
        self._lib = load_library()
        self.n, self.mc, self.nb = int(nbVar), int(nbCstr), (int(nbVar) if useBounds else 0)
        self.m = self.mc + self.nb
        self.capacity = int(batch_capacity)
        self._h = None
        self._mh = C.c_void_p()
        if devices is not None:
            arr = (C.c_int32 * len(devices))(*[int(v) for v in devices])
            rc = self._lib.jrlqp_multi_create(C.byref(self._mh), self.n, self.mc, int(bool(useBounds)), self.capacity, arr, len(devices))
        else:
            rc = self._lib.jrlqp_multi_create(C.byref(self._mh), self.n, self.mc, int(bool(useBounds)), self.capacity, None, int(n_devices))
        if rc != 0:
            msg = self._lib.jrlqp_multi_last_error(self._mh).decode() if self._mh else "allocation failed"
            if self._mh:
                self._lib.jrlqp_multi_destroy(self._mh)
                self._mh = None
            raise JrlQpError(f"jrlqp_multi_create failed ({rc}): {msg}")
        self.n_devices = int(self._lib.jrlqp_multi_device_count(self._mh))
        self.devices = [int(self._lib.jrlqp_multi_device(self._mh, k)) for k in range(self.n_devices)]
        self._h = self._lib.jrlqp_multi_solver(self._mh, 0)  # introspection (kernel_info, host_g_bytes) goes to the first device's solver
        self._options = SolverOptions()
        self.last = None

    def __del__(self):
        mh = getattr(self, "_mh", None)
        self._h = None
        if mh:
            self._lib.jrlqp_multi_destroy(mh)
            self._mh = None

    def options(self, opt=None):
        if opt is None:
            return self._options
        self._options = opt
        o = _Options(opt.maxIter_, opt.bigBnd_, int(opt.warmStart_), opt.logFlags_)
        if self._lib.jrlqp_multi_set_options(self._mh, C.byref(o)) != 0:
            raise JrlQpError("jrlqp_multi_set_options failed")
        return self

    def shard(self, batch, k):
        lo, hi = C.c_int64(), C.c_int64()
        if self._lib.jrlqp_multi_shard(self._mh, int(batch), int(k), C.byref(lo), C.byref(hi)) != 0:
            raise JrlQpError("jrlqp_multi_shard failed")
        return int(lo.value), int(hi.value)

    def set_balancing(self, on):
        """Shares of the shards follow the measured per-device throughput of the previous calls (default) or stay equal."""
        if self._lib.jrlqp_multi_set_balancing(self._mh, int(bool(on))) != 0:
            raise JrlQpError("jrlqp_multi_set_balancing failed")
        return self

    def weights(self):
        w = (C.c_double * self.n_devices)()
        self._lib.jrlqp_multi_get_weights(self._mh, w)
        return [float(v) for v in w]

    def _host_call(self, pb, res, experimental):
        fn = self._lib.jrlqp_multi_solve_batch_warm_host if experimental else self._lib.jrlqp_multi_solve_batch_host
        return fn(self._mh, C.byref(pb), C.byref(res))

    def _last_error(self):
        return self._lib.jrlqp_multi_last_error(self._mh).decode()

    def solve_device(self, *a, **k):
        raise JrlQpError("the multi-GPU handle takes host arrays (solve); device pointers belong to one device (BatchedGoldfarbIdnaniSolver)")

    solve_sequence = solve_sequence_device = solve_device


class GoldfarbIdnaniSolver:
    """One-QP-per-call mirror of jrl::qp::GoldfarbIdnaniSolver (batch of 1 through the same kernels).

    G is n x n (symmetric, lower triangle read); C is n x nbCstr with one constraint per COLUMN, as
    in the reference (tests pass `qpp.C.transpose()`, tests/GoldfarbIdnaniSolverTest.cpp:89).
    After solve(), G's lower triangle holds the Cholesky factor, as with the reference.
    """

    def __init__(self, nbVar=0, nbCstr=0, useBounds=False, device=0):
        self._device = device
        self._impl = None
        self._opt = SolverOptions()
        self._res = None
        if nbVar > 0:
            self.resize(nbVar, nbCstr, useBounds)

    def resize(self, nbVar, nbCstr, useBounds):
        if self._impl is None or (self._impl.n, self._impl.mc, self._impl.nb > 0) != (nbVar, nbCstr, bool(useBounds)):
            self._impl = BatchedGoldfarbIdnaniSolver(nbVar, nbCstr, useBounds, 1, self._device)
            self._impl.options(self._opt)

    def options(self, opt=None):
        if opt is None:
            return self._opt
        self._opt = opt
        if self._impl is not None:
            self._impl.options(opt)
        return self

    _experimental = False

    def solve(self, G, a, Cmat, bl, bu, xl, xu, as_=None):
        G = np.asarray(G)
        n = G.shape[0]
        Cmat = np.asarray(Cmat, dtype=np.float64).reshape(n, -1)
        nbCstr = Cmat.shape[1]
        useBnd = xl is not None and np.size(xl) > 0
        self.resize(n, nbCstr, useBnd)
        # column-major n x mc  ==  row-major [mc, n]
        Cm = np.ascontiguousarray(Cmat.T)
        Gc = np.ascontiguousarray(np.asarray(G, dtype=np.float64).T)  # column-major blocks
        st = self._impl.solve(Gc[None], np.asarray(a, dtype=np.float64)[None], Cm[None] if nbCstr else None,
                              np.asarray(bl, dtype=np.float64)[None] if nbCstr else None,
                              np.asarray(bu, dtype=np.float64)[None] if nbCstr else None,
                              np.asarray(xl, dtype=np.float64)[None] if useBnd else None,
                              np.asarray(xu, dtype=np.float64)[None] if useBnd else None, want_L=True,
                              **self._extra(as_, nbCstr + (n if useBnd else 0)))
        self._res = self._impl.last
        if st != TerminationStatus.NON_POS_HESSIAN and isinstance(G, np.ndarray) and G.dtype == np.float64 and G.flags.writeable:
            Lf = self._res["L"][0].T  # back to (i, j) indexing
            il = np.tril_indices(n)
            G[il] = Lf[il]  # G is an in/out argument in the reference (src/GoldfarbIdnaniSolver.cpp:58)
        return st

    def _extra(self, as_, m):
        if as_ is not None and len(as_) > 0:
            raise JrlQpError("the stable GoldfarbIdnaniSolver takes no active-set guess; use experimental.GoldfarbIdnaniSolver")
        return {}

    def solution(self):
        return self._res["x"][0]

    def multipliers(self):
        return self._res["u"][0]

    def objectiveValue(self):
        return float(self._res["f"][0])

    def iterations(self):
        return int(self._res["iterations"][0])

    def activeSet(self):
        return [ActivationStatus(int(v)) for v in self._res["active_set"][0]]

    def resetActiveSet(self):
        pass  # the stable solver resets its active set at every solve (src/GoldfarbIdnaniSolver.cpp:75)


class ExperimentalGoldfarbIdnaniSolver(GoldfarbIdnaniSolver):
    """Mirror of jrl::qp::experimental::GoldfarbIdnaniSolver (include/jrl-qp/experimental/GoldfarbIdnaniSolver.h:15-39):
    solve(G, a, C, bl, bu, xl, xu, as=[]) with SolverOptions::warmStart. With warm start on and an empty
    `as`, the active set of the previous solve is reused (src/experimental/GoldfarbIdnaniSolver.cpp:58-61)."""

    def _extra(self, as_, m):
        if as_ is not None and len(as_) > 0:
            if not self._opt.warmStart_:
                raise JrlQpError("Non-empty active set used with cold start option.")  # the reference asserts
            guess = np.array([int(v) for v in as_], dtype=np.int8)
        elif self._opt.warmStart_ and self._res is not None and self._res["active_set"].shape[1] == m:
            guess = self._res["active_set"][0].copy()
        else:
            guess = None
        return {"experimental": True, "as_in": None if guess is None else guess[None]}

    def resetActiveSet(self):
        self._res = None  # DualSolver::resetActiveSet (src/DualSolver.cpp:85-88)


class experimental:  # namespace jrl::qp::experimental
    GoldfarbIdnaniSolver = ExperimentalGoldfarbIdnaniSolver
