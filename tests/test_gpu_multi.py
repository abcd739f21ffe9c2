"""Multi-GPU entry of the C-ABI (jrlqp_multi_*): one host batch scattered over the GPUs of the box by contiguous
shards and gathered in place, compared with the ORACLE bit for bit. The sharded path is exercised on
every visible device; with one GPU the handle is also created over [0, 0] (two shards, two host threads, two solvers
on the same device), so the scatter / gather logic is covered by the single-GPU tier as well."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import jrl_qp_b200  # noqa: F401,E402
from jrl_qp_b200 import problems as P, sharding, solver as S  # noqa: E402

pytestmark = pytest.mark.gpu


def _ndev():
    import torch
    return torch.cuda.device_count()


def _check(sv, pb, experimental=False, as_in=None, warm=False):
    import pyoracle as po
    kw = {}
    if experimental:
        kw = dict(experimental=True, warm_start=warm, as_in=as_in)
        sv.options(S.SolverOptions().warmStart(warm))
    ref = po.solve_batch(pb.G, pb.a, pb.C, pb.bl, pb.bu, pb.xl, pb.xu, nthreads=os.cpu_count(), **kw)
    sv.solve(pb.G, pb.a, pb.C, pb.bl, pb.bu, pb.xl, pb.xu, experimental=experimental, as_in=as_in)
    g = sv.last
    for k in ("x", "u", "f", "iterations", "status", "active_set", "n_active", "active_list"):
        assert np.array_equal(g[k], ref[k]), k
    return ref


@pytest.mark.parametrize("devices", [None, [0, 0], [0, 0, 0]])
def test_multi_entry_equals_oracle(devices):
    if devices is None and _ndev() < 2:
        pytest.skip("needs 2 GPUs (the same logic runs below over [0, 0])")
    B = 1003  # not a multiple of the number of shards
    pb = P.random_problems(P.config_A(), B, seed=4242)
    sv = S.MultiGpuGoldfarbIdnaniSolver(pb.n, pb.mc, True, B, devices=devices)
    g = sv.n_devices
    assert g >= 2
    # equal shares: the shards of the library are those of jrl-qp_b200/sharding.py
    assert [sv.shard(B, k) for k in range(g)] == sharding.all_shards(B, g)
    sv.set_balancing(False)
    _check(sv, pb)
    assert [sv.shard(B, k) for k in range(g)] == sharding.all_shards(B, g)
    # a smaller batch on the same handle, then a larger one again (capacity-managed staging)
    _check(sv, pb.slice(0, 5))
    _check(sv, pb.slice(3, 700))


def test_multi_entry_shared_arrays_and_warm_start():
    B = 257
    devices = None if _ndev() >= 2 else [0, 0]
    pb = P.random_problems(P.config_B(), B, seed=99)
    # C and the bounds shared by the batch (stride 0): every device gets them once
    shared = P.ProblemBatch(pb.G, pb.a, pb.C[0], pb.bl[0], pb.bu[0], pb.xl[0], pb.xu[0])
    sv = S.MultiGpuGoldfarbIdnaniSolver(pb.n, pb.mc, True, B, devices=devices)
    ref = _check(sv, shared)
    # experimental solver, warm-started from the exact active set: zero iterations (tests/GoldfarbIdnaniSolverTest.cpp:176-181)
    _check(sv, shared, experimental=True, as_in=ref["active_set"], warm=True)
    assert (sv.last["iterations"][ref["status"] == 0] == 0).all()


def test_multi_entry_load_balancing_keeps_the_bits():
    """Shares that follow the measured throughput: contiguous shards that still cover the batch, same results."""
    B = 3 * 4096 + 17
    devices = None if _ndev() >= 2 else [0, 0, 0]
    pb = P.random_problems(P.config_B(), B, seed=5)
    sv = S.MultiGpuGoldfarbIdnaniSolver(pb.n, pb.mc, True, B, devices=devices)
    g = sv.n_devices
    for _ in range(4):
        _check(sv, pb)
        sh = [sv.shard(B, k) for k in range(g)]
        assert sh[0][0] == 0 and sh[-1][1] == B and all(sh[k][1] == sh[k + 1][0] for k in range(g - 1))
        assert all(hi - lo <= -(-B // g) * 3 // 2 + 1 for lo, hi in sh)
        w = sv.weights()
        assert abs(sum(w) - 1.0) < 1e-12 and max(w) <= 1.5 / g + 1e-12


def test_multi_entry_errors():
    sv = S.MultiGpuGoldfarbIdnaniSolver(5, 3, True, 8, devices=[0, 0])
    pb = P.random_problems(P.ProblemCharacteristics(5, 1, 2, 1, 0, 1, 0, True, False), 16, seed=1)
    with pytest.raises(S.JrlQpError):  # batch above the capacity of the handle: JRLQP_ERR_CAPACITY
        sv.solve(pb.G, pb.a, pb.C, pb.bl, pb.bu, pb.xl, pb.xu)
    with pytest.raises(S.JrlQpError):
        S.MultiGpuGoldfarbIdnaniSolver(5, 3, True, 8, devices=[0, 99])


def test_host_link_probe_runs():
    agg, per = S.measure_host_link(1, nbytes=64 << 20, reps=2, direction=0)
    assert agg > 1.0 and len(per) == 1
