"""GPU parity tests (-m gpu) of the structured solver (experimental::BlockGISolver, batched): the CUDA kernel, called
through the C-ABI (jrlqp_blockgi_*), against the CPU oracle (oracle/block_oracle.cpp) on the same seeded inputs — status,
iteration count, active set, ordered active list, x, multipliers and objective bit for bit — and against the reference's
own acceptance criterion (same status as the dense solver, solution within 1e-8: tests/BlockGISolverTest.in.cpp:120-122,
167-169, 216-217, 304-305)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import block_cases as bc
import pyoracle as po
import jrl_qp_b200  # noqa: F401
from jrl_qp_b200 import solver as S
from jrl_qp_b200.blockgi import BatchedBlockGISolver, BlockGISolver, StructuredC
from jrl_qp_b200.structured import Type

pytestmark = pytest.mark.gpu
ALL_TYPES = [Type.TriBlockDiagonal, Type.BlockArrowDown, Type.BlockArrowUp]
KEYS = ("status", "iterations", "n_active", "active_set", "active_list", "x", "u", "f")


def _both(pb, max_iter=None, capacity=None):
    B = pb.a.shape[0] if pb.a.ndim == 2 else 1
    sv = BatchedBlockGISolver(pb.stG, pb.stC, pb.xl is not None, capacity or B)
    kw = {}
    if max_iter is not None:
        sv.options(S.SolverOptions().maxIter(max_iter))
        kw["max_iter"] = max_iter
    before = S.launch_count()
    sv.solve(pb.Gdata, pb.a, pb.Cdata, pb.bl, pb.bu, pb.xl, pb.xu)
    assert S.launch_count() >= before + 2, "factorisation + solver kernels were not launched"
    ref = po.block_solve_batch(pb.stG, pb.stC, pb.Gdata, pb.a, pb.Cdata, pb.bl, pb.bu, pb.xl, pb.xu, nthreads=os.cpu_count(), **kw)
    return sv, sv.last, ref


def _assert_bit_exact(g, ref):
    for k in KEYS:
        assert np.array_equal(g[k], ref[k]), f"{k} differs from the oracle"


@pytest.mark.parametrize("layout", ["packed", "dense"])
@pytest.mark.parametrize("type", ALL_TYPES)
def test_reference_test_sizes(type, layout):
    # tests/BlockGISolverTest.in.cpp:68-170
    pb = bc.random_block_problem(type, [3, 5, 2, 3], [3, 3, 3, 3], 96, seed=31 + int(type), layout=layout, shift=0.05)
    sv, g, ref = _both(pb)
    _assert_bit_exact(g, ref)
    assert (g["status"] == 0).all() and g["iterations"].max() > 0
    rd = bc.dense_solution(pb)
    assert np.array_equal(g["status"], rd["status"])
    for k in range(96):
        assert bc.is_approx(g["x"][k], rd["x"][k], 1e-8)
        assert bc.is_approx(g["x"][k], pb.x_planted[k], 1e-6)


@pytest.mark.parametrize("type,sizes,mi,batch,bounds", [
    (Type.TriBlockDiagonal, [12] * 32, [12] * 32, 48, False),   # config E (BASELINE.json configs[4])
    (Type.BlockArrowDown, [12] * 8, [12] * 8, 64, True),
    (Type.BlockArrowUp, [12] * 8, [6] * 8, 64, True),
    (Type.TriBlockDiagonal, [43] * 4, [20] * 4, 16, False),     # two warps per QP
    (Type.BlockArrowUp, [42] * 3, [5] * 3, 16, True),
    (Type.TriBlockDiagonal, [7, 1, 30, 2, 70, 5], [3, 1, 9, 0, 20, 2], 16, True),   # four warps, an empty C block
    (Type.BlockArrowDown, [5, 33, 1, 9], [4, 12, 1, 5], 32, False),
    (Type.TriBlockDiagonal, [6], [9], 32, True),
])
def test_bit_exact_vs_oracle(type, sizes, mi, batch, bounds):
    pb = bc.random_block_problem(type, sizes, mi, batch, seed=101 + len(sizes), bounds=bounds, shift=0.05)
    sv, g, ref = _both(pb)
    _assert_bit_exact(g, ref)
    assert (g["status"] == 0).all()
    for k in range(batch):
        assert bc.is_approx(g["x"][k], pb.x_planted[k], 1e-6)


@pytest.mark.parametrize("fast", [0, 1, 2, 4, 7])
@pytest.mark.parametrize("sizes,mi,batch,bounds", [
    ([12] * 32, [12] * 32, 24, False),  # config E: aligned hints, full blocks only
    ([12] * 6, [7] * 6, 32, True),      # bounds: hints in the middle of a tile (partial first block, exact links)
    ([8] * 9, [5] * 9, 32, True),
    ([16] * 5, [9] * 5, 32, True),      # padded stage columns
])
def test_fast_paths_bit_exact(sizes, mi, batch, bounds, fast, monkeypatch):
    """Round 2 fast paths of the kernel (blockgi.cuh BGF_*): warp-level structured solves fed by TMA bulk copies, the
    orthonormal sequence on a register-resident vector, the blocked R solve - each alone and all together, against the
    oracle bit for bit (JRLQP_BLOCKGI_FAST is read when the solver is created)."""
    monkeypatch.setenv("JRLQP_BLOCKGI_FAST", str(fast))
    pb = bc.random_block_problem(Type.TriBlockDiagonal, sizes, mi, batch, seed=211 + len(sizes), bounds=bounds, shift=0.05)
    sv, g, ref = _both(pb)
    _assert_bit_exact(g, ref)
    assert (g["status"] == 0).all() and g["iterations"].max() > 0


def test_fast_paths_more_than_512_variables():
    # 128 classes x 8 entries per thread in the passes over Q (n > 512), uniform tiles
    pb = bc.random_block_problem(Type.TriBlockDiagonal, [12] * 44, [4] * 44, 4, seed=5, shift=0.05)
    sv, g, ref = _both(pb)
    _assert_bit_exact(g, ref)
    assert (g["status"] == 0).all()


def test_fast_paths_unaligned_instances_fall_back():
    # an odd stride between instances: no bulk copies, the general solves run (same bits)
    torch = pytest.importorskip("torch")
    pb = bc.random_block_problem(Type.TriBlockDiagonal, [12] * 5, [6] * 5, 8, seed=17, shift=0.05)
    B, n = 8, pb.n
    dev = torch.device("cuda:0")
    stride = pb.Gdata.shape[1] + 1
    Gp = np.zeros((B, stride))
    Gp[:, :-1] = pb.Gdata
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    G, a, Cd, bl, bu = t(Gp), t(pb.a), t(pb.Cdata), t(pb.bl), t(pb.bu)
    x = torch.empty((B, n), dtype=torch.float64, device=dev)
    status = torch.empty(B, dtype=torch.int32, device=dev)
    sv = BatchedBlockGISolver(pb.stG, pb.stC, False, B)
    sv.solve_device(B, G, a, Cd, bl, bu, None, None, x, status=status, G_stride=stride)
    torch.cuda.synchronize()
    ref = po.block_solve_batch(pb.stG, pb.stC, pb.Gdata, pb.a, pb.Cdata, pb.bl, pb.bu)
    assert np.array_equal(x.cpu().numpy(), ref["x"]) and (status.cpu().numpy() == 0).all()


def test_unshifted_ill_conditioned():
    pb = bc.random_block_problem(Type.TriBlockDiagonal, [3, 5, 2, 3], [3, 3, 3, 3], 128, seed=77)
    sv, g, ref = _both(pb)
    _assert_bit_exact(g, ref)


def test_multiik_sequential_shared_G_and_C():
    pb, d = bc.multiik_sequential(24, scale=1e-3)
    sv, g, ref = _both(pb)
    _assert_bit_exact(g, ref)
    assert (g["status"] == 0).all()
    rd = po.solve_batch(pb.Gdense, pb.a, pb.Cdense, pb.bl, pb.bu, nthreads=os.cpu_count())
    assert np.abs(d["sol"] - g["x"][0]).max() <= 1e-4
    for k in range(24):
        assert bc.is_approx(g["x"][k], rd["x"][k], 1e-8)


def test_multiik_simultaneous_shared_G_and_C():
    pb, d = bc.multiik_simultaneous(48, scale=1e-3)
    sv, g, ref = _both(pb)
    _assert_bit_exact(g, ref)
    assert (g["status"] == 0).all()
    rd = po.solve_batch(pb.Gdense, pb.a, pb.Cdense, pb.bl, pb.bu, pb.xl, pb.xu, nthreads=os.cpu_count())
    for k in range(48):
        assert bc.is_approx(g["x"][k], rd["x"][k], 1e-8)


def test_status_codes_and_iteration_cap():
    pb = bc.random_block_problem(Type.TriBlockDiagonal, [3, 4, 3], [2, 2, 2], 6, seed=1, shift=0.05)
    pb.bu = pb.bu.copy()
    pb.bu[1, 2] = pb.bl[1, 2]  # an equality: INCONSISTENT_INPUT
    pb.Gdata = pb.Gdata.copy()
    pb.Gdata[2, pb.stG.diag_offset[1]] = -1.0  # NON_POS_HESSIAN
    Cd = pb.Cdense.copy()
    Cd[3, 1] = Cd[3, 0]  # two parallel constraints with disjoint slabs: INFEASIBLE
    pb.bl, pb.bu = pb.bl.copy(), pb.bu.copy()
    pb.bl[3, 0], pb.bu[3, 0] = 0.0, 1.0
    pb.bl[3, 1], pb.bu[3, 1] = 2.0, 3.0
    pb.bl[4, 3], pb.bu[4, 3] = pb.bu[4, 3] + 1.0, pb.bl[4, 3] - 1.0  # bl > bu: sequential scan replay
    pb.Cdata = pb.stC.pack(Cd)
    sv, g, ref = _both(pb)
    _assert_bit_exact(g, ref)
    assert list(g["status"][:4]) == [0, 1, 2, 3]
    sv, g, ref = _both(pb, max_iter=2)
    _assert_bit_exact(g, ref)
    assert (g["status"] == 4).any()


def test_device_entry_factorises_G_in_place():
    torch = pytest.importorskip("torch")
    pb = bc.random_block_problem(Type.BlockArrowUp, [5, 7, 4], [4, 4, 4], 32, seed=9, shift=0.05)
    B, n, mc = 32, pb.n, pb.mc
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    G, a, Cd, bl, bu = t(pb.Gdata), t(pb.a), t(pb.Cdata), t(pb.bl), t(pb.bu)
    x = torch.empty((B, n), dtype=torch.float64, device=dev)
    status = torch.empty(B, dtype=torch.int32, device=dev)
    sv = BatchedBlockGISolver(pb.stG, pb.stC, False, B)
    sv.solve_device(B, G, a, Cd, bl, bu, None, None, x, status=status)
    torch.cuda.synchronize()
    ref = po.block_solve_batch(pb.stG, pb.stC, pb.Gdata, pb.a, pb.Cdata, pb.bl, pb.bu, want_L=True)
    assert np.array_equal(x.cpu().numpy(), ref["x"]) and (status.cpu().numpy() == 0).all()
    assert np.array_equal(G.cpu().numpy(), ref["L"]), "G does not hold the oracle's factor"
    # shared G: left untouched
    G1 = t(pb.Gdata[0])
    sv.solve_device(B, G1, a, Cd, bl, bu, None, None, x, status=status, shared=("G",))
    torch.cuda.synchronize()
    assert np.array_equal(G1.cpu().numpy(), pb.Gdata[0])
    ref1 = po.block_solve_batch(pb.stG, pb.stC, pb.Gdata[0], pb.a, pb.Cdata, pb.bl, pb.bu)
    assert np.array_equal(x.cpu().numpy(), ref1["x"])


def test_single_problem_mirror_reads_like_the_reference_test():
    # tests/BlockGISolverTest.in.cpp:68-123
    pb = bc.random_block_problem(Type.TriBlockDiagonal, [3, 5, 2, 3], [3, 3, 3, 3], 1, seed=5, shift=0.05)
    G = (pb.stG, pb.Gdata[0])
    Cs = StructuredC(pb.stC, pb.Cdata[0])
    solverB = BlockGISolver(13, 12, False)
    retB = solverB.solve(G, pb.a[0], Cs, pb.bl[0], pb.bu[0], np.zeros(0), np.zeros(0))
    rd = bc.dense_solution(pb)
    assert int(retB) == int(rd["status"][0]) == 0
    assert bc.is_approx(solverB.solution(), rd["x"][0], 1e-8)
    assert solverB.iterations() > 0 and len(solverB.activeSet()) == 12
    assert abs(solverB.objectiveValue() - rd["f"][0]) <= 1e-9 * max(1.0, abs(rd["f"][0]))


def test_argument_errors():
    pb = bc.random_block_problem(Type.TriBlockDiagonal, [3, 4], [2, 2], 4, seed=2)
    sv = BatchedBlockGISolver(pb.stG, pb.stC, False, 2)
    with pytest.raises(S.JrlQpError):
        sv.solve(pb.Gdata, pb.a, pb.Cdata, pb.bl, pb.bu)  # batch 4 > capacity 2
    from jrl_qp_b200.structured import CStructure
    with pytest.raises(S.JrlQpError):
        BatchedBlockGISolver(pb.stG, CStructure.packed([3, 5], [2, 2]), False, 2)  # C does not cover the variables
