"""Robustness of the handle-level behaviour (round-1 advisor findings) and the reference's no-heap contract:
  * several live solvers that share a kernel instantiation but not a shared-memory size;
  * calls of one handle on different streams (serialised on the device);
  * more than nbVar equalities on the COLD path: a status, not a corrupted batch;
  * NaN solutions fail the batch verifier;
  * mis-shaped host arrays are refused before any copy;
  * a re-used handle allocates nothing on its second solve (tests/GoldfarbIdnaniSolverTest.cpp:101-125,
    src/internal/memoryChecks.cpp:19-23: the reference's EIGEN_RUNTIME_NO_MALLOC check, restated for device memory)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import jrl_qp_b200  # noqa: F401,E402
from jrl_qp_b200 import problems as P, solver as S  # noqa: E402

pytestmark = pytest.mark.gpu


def _po():
    import pyoracle as po
    return po


def _oracle(pb, **kw):
    return _po().solve_batch(pb.G, pb.a, pb.C, pb.bl, pb.bu, pb.xl, pb.xu, nthreads=os.cpu_count(), **kw)


def _same(g, ref):
    for k in ("status", "iterations", "n_active", "active_set", "active_list", "x", "u", "f"):
        assert np.array_equal(g[k], ref[k]), k


def test_interleaved_live_solvers_of_different_sizes_in_one_kernel_instantiation():
    """n = 128 and n = 100 both run gi_dense_cta_kernel<4, ...>: the later, smaller handle must not lower the
    dynamic shared-memory limit the earlier one needs (and likewise n = 50 / n = 40 at <2, ...>)."""
    for big, small in ((128, 100), (50, 40)):
        chs = [P.ProblemCharacteristics(nn, nn // 5, nn // 2, nn // 8, 0, nn // 10, 0, True, False) for nn in (big, small)]
        pbs = [P.random_problems(ch, 24, seed=7 + ch.nVar) for ch in chs]
        svs = [S.BatchedGoldfarbIdnaniSolver(pb.n, pb.mc, True, 24) for pb in pbs]  # both alive, the small one created last
        refs = [_oracle(pb) for pb in pbs]
        for _ in range(2):
            for sv, pb, ref in zip(svs, pbs, refs):  # big, small, big, small
                sv.solve(pb.G, pb.a, pb.C, pb.bl, pb.bu, pb.xl, pb.xu)
                _same(sv.last, ref)


def test_one_handle_on_two_streams_is_serialised():
    import torch
    dev = torch.device("cuda", 0)
    pb = P.random_problems(P.config_A(), 2048, seed=3)
    ref = _oracle(pb)
    d = {k: torch.from_numpy(getattr(pb, k)).to(dev) for k in ("G", "a", "C", "bl", "bu", "xl", "xu")}
    sv = S.BatchedGoldfarbIdnaniSolver(pb.n, pb.mc, True, pb.batch)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    outs = []
    for st in (s1, s2, s1, s2):
        x = torch.empty((pb.batch, pb.n), dtype=torch.float64, device=dev)
        it = torch.empty(pb.batch, dtype=torch.int32, device=dev)
        sv.solve_device(pb.batch, d["G"], d["a"], d["C"], d["bl"], d["bu"], d["xl"], d["xu"], x, iterations=it, stream=st.cuda_stream)
        outs.append((x, it))
    torch.cuda.synchronize()
    for x, it in outs:
        assert np.array_equal(x.cpu().numpy(), ref["x"]) and np.array_equal(it.cpu().numpy(), ref["iterations"])


@pytest.mark.parametrize("path", [1, 2])
def test_cold_path_with_more_than_n_equalities_reports_overconstrained(path):
    n, mc, B = 6, 9, 5
    rng = np.random.default_rng(1)
    A = rng.normal(size=(B, n, n))
    G = A @ A.transpose(0, 2, 1) + np.eye(n)
    a = rng.normal(size=(B, n))
    Cm = rng.normal(size=(B, mc, n))
    bl = rng.normal(size=(B, mc))
    bu = bl + 1.0
    bu[1] = bl[1]            # instance 1: nine equalities on six variables
    bu[3, :7] = bl[3, :7]    # instance 3: seven
    bu[4, :6] = bl[4, :6]    # instance 4: exactly n (fine)
    sv = S.BatchedGoldfarbIdnaniSolver(n, mc, False, B).set_kernel_path(path)
    sv.solve(G, a, Cm, bl, bu)
    ref = _po().solve_batch(G, a, Cm, bl, bu)
    _same(sv.last, ref)
    st = sv.last["status"]
    assert st[1] == st[3] == S.TerminationStatus.OVERCONSTRAINED_PROBLEM
    assert all(st[k] in (0, 3) for k in (0, 2, 4))  # solved or INFEASIBLE (random slabs), never the overflow status


def test_nan_solution_fails_the_batch_verifier():
    # no constraints and no bounds: stationarity is the only test, so a NaN dropped by the reductions would pass everything
    pb = P.random_problems(P.ProblemCharacteristics(8), 4, seed=2)
    sv = S.BatchedGoldfarbIdnaniSolver(pb.n, 0, False, pb.batch)
    sv.solve(pb.G, pb.a, None, None, None)
    x, u = sv.last["x"].copy(), sv.last["u"].copy()
    flags, _, nfail = sv.test_kkt(x, u, pb.G, pb.a, None, None, None)
    assert nfail == 0 and (flags == 3).all()
    x[2, 3] = np.nan
    flags, resid, nfail = sv.test_kkt(x, u, pb.G, pb.a, None, None, None)
    assert nfail == 1 and (flags[2] & 1) == 0 and (np.delete(flags, 2) == 3).all()
    o = _po().kkt_check_batch(x, u, pb.G, pb.a, None, None, None)
    assert np.array_equal(flags, o[0]) and o[2] == 1


def test_misshaped_arrays_are_refused():
    pb = P.random_problems(P.config_B(), 8, seed=1)
    sv = S.BatchedGoldfarbIdnaniSolver(pb.n, pb.mc, True, 8)
    with pytest.raises(S.JrlQpError):
        sv.solve(pb.G[:, :-1], pb.a, pb.C, pb.bl, pb.bu, pb.xl, pb.xu)
    with pytest.raises(S.JrlQpError):
        sv.solve(pb.G, pb.a, pb.C[:, :-1], pb.bl, pb.bu, pb.xl, pb.xu)
    with pytest.raises(S.JrlQpError):
        sv.solve(pb.G, pb.a[:4], pb.C, pb.bl, pb.bu, pb.xl, pb.xu)
    with pytest.raises(S.JrlQpError):
        sv.solve(pb.G, pb.a, pb.C, pb.bl, pb.bu, None, None)
    sv.solve(pb.G, pb.a, pb.C, pb.bl, pb.bu, pb.xl, pb.xu)
    assert (sv.last["status"] == 0).all()


@pytest.mark.parametrize("cfg", ["A", "B", "warm"])
def test_reused_handle_allocates_no_device_memory(cfg):
    """The reference runs its second solve under EIGEN_RUNTIME_NO_MALLOC; here: the free device memory does not
    move across the second (and third) solve of a handle, host entry point included (its staging is sized at the
    first call by the capacity of the handle), and neither does the number of live allocations reported by the
    CUDA memory pools."""
    import torch
    ch = P.config_B() if cfg != "A" else P.config_A()
    pb = P.random_problems(ch, 3000, seed=11)
    sv = S.BatchedGoldfarbIdnaniSolver(pb.n, pb.mc, True, 4096)
    kw = {}
    if cfg == "warm":
        sv.options(S.SolverOptions().warmStart(True))
        sv.solve(pb.G, pb.a, pb.C, pb.bl, pb.bu, pb.xl, pb.xu)
        kw = dict(experimental=True, as_in=sv.last["active_set"].copy())
    sv.solve(pb.G, pb.a, pb.C, pb.bl, pb.bu, pb.xl, pb.xu, **kw)  # first use: staging is created
    first = {k: v.copy() for k, v in sv.last.items() if isinstance(v, np.ndarray)}
    torch.cuda.synchronize()
    free0, _ = torch.cuda.mem_get_info(0)
    for B in (3000, 1000, 4096):
        sub = pb.slice(0, min(B, pb.batch))
        kw2 = dict(kw)
        if "as_in" in kw2:
            kw2["as_in"] = kw2["as_in"][:sub.batch]
        sv.solve(sub.G, sub.a, sub.C, sub.bl, sub.bu, sub.xl, sub.xu, **kw2)
        torch.cuda.synchronize()
        free1, _ = torch.cuda.mem_get_info(0)
        assert free1 == free0, f"device memory moved by {free0 - free1} bytes on a re-used handle (batch {B})"
    sv.solve(pb.G, pb.a, pb.C, pb.bl, pb.bu, pb.xl, pb.xu, **kw)
    for k, v in first.items():
        assert np.array_equal(sv.last[k], v), k
