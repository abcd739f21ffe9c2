"""GPU parity tests (-m gpu) of the structured decompositions: the CUDA kernels, called through the
C-ABI (jrlqp_structured_*), against the CPU oracle (oracle/decomp_oracle.cpp) on the same seeded
inputs — bit for bit — and against the reference's own acceptance criterion (agreement with the dense
Cholesky factor / dense triangular solves to 1e-8, tests/triBlockDiagLLTTest.cpp:50,
tests/blockArrowLLTTest.cpp:55)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import pyoracle as po
import structured_cases as sc
import jrl_qp_b200  # noqa: F401
from jrl_qp_b200 import solver as S
from jrl_qp_b200.structured import Structure, StructuredG, Type

pytestmark = pytest.mark.gpu
ALL_TYPES = [Type.TriBlockDiagonal, Type.BlockArrowDown, Type.BlockArrowUp]


def _factor_both(st, H, kernel=0):
    data_ref = st.pack(H)
    ok_ref = po.decomp_llt(st, data_ref, nthreads=os.cpu_count())
    g = StructuredG(st, st.pack(H))
    if kernel:
        g.set_kernel(kernel)
    before = S.launch_count()
    ok = g.lltInPlace()
    assert S.launch_count() > before, "no CUDA kernel was launched"
    return g, ok, data_ref, ok_ref


@pytest.mark.parametrize("layout", ["dense", "packed"])
@pytest.mark.parametrize("type", ALL_TYPES)
def test_reference_test_sizes(type, layout):
    sizes = [3, 5, 2, 3]
    st = Structure.dense(type, sizes) if layout == "dense" else Structure.packed(type, sizes)
    H = sc.make_H(type, sizes, 64, seed=21)
    g, ok, data_ref, ok_ref = _factor_both(st, H)
    assert ok.all() and ok_ref.all()
    assert np.array_equal(g.data, data_ref), "factor not bit-identical to the oracle"
    L = sc.factor_from_data(st, g.data)
    Lref = sc.dense_factor(type, sizes, H)
    assert np.abs(L - Lref).max() <= 1e-8 * np.abs(Lref).max()


@pytest.mark.parametrize("type,sizes,batch", [
    (Type.TriBlockDiagonal, [12] * 32, 512),       # config E (BASELINE.json configs[4])
    (Type.BlockArrowDown, [12] * 32, 256),
    (Type.BlockArrowUp, [12] * 32, 256),
    (Type.TriBlockDiagonal, [43] * 9, 64),         # MultiIK sequential structure
    (Type.BlockArrowUp, [42] * 5, 64),             # MultiIK simultaneous structure
    (Type.TriBlockDiagonal, [7, 1, 30, 2, 64, 5], 32),
    (Type.BlockArrowDown, [5, 33, 1, 9], 32),
    (Type.BlockArrowUp, [9, 33, 1, 70], 32),
    (Type.TriBlockDiagonal, [6], 8),
    (Type.BlockArrowDown, [70, 96], 4),
])
def test_llt_and_solves_bit_exact(type, sizes, batch):
    st = Structure.packed(type, sizes)
    H = sc.make_H(type, sizes, batch, seed=33, shift=1.0)
    g, ok, data_ref, ok_ref = _factor_both(st, H)
    assert np.array_equal(ok, ok_ref.astype(bool)) and ok.all()
    assert np.array_equal(g.data, data_ref), "factor not bit-identical to the oracle"
    rng = np.random.default_rng(2)
    n = st.n
    for ncols in (1, 3):
        V = rng.uniform(-1, 1, (batch, ncols, n))
        for transpose in (False, True):
            ref = po.decomp_solve(st, data_ref, V.copy(), transpose=transpose, nthreads=os.cpu_count())
            out = g.solveLTranspose(V.copy()) if transpose else g.solveL(V.copy())
            assert np.array_equal(out, ref), f"solve (transpose={transpose}) not bit-identical to the oracle"
    # L then L^T = H^-1 (reference criterion, tests/blockArrowLLTTest.cpp:162-171)
    v = rng.uniform(-1, 1, (batch, n))
    x = g.solveInPlaceLTranspose(g.solveL(v.copy()))
    ref = np.linalg.solve(H, v[:, :, None])[:, :, 0]
    assert np.abs(x - ref).max() <= 1e-8 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("type", ALL_TYPES)
def test_hint_windows(type):
    """every (start, end) window of tests/blockArrowLLTTest.cpp:57-98, one instance per window"""
    sizes = [3, 5, 2, 3]
    n = 13
    st = Structure.packed(type, sizes)
    H = sc.make_H(type, sizes, 1, seed=5)
    data = st.pack(H)
    assert po.decomp_llt(st, data).all()
    g = StructuredG(st, st.pack(H))
    assert g.lltInPlace().all()
    rng = np.random.default_rng(9)
    for i in range(n):
        for j in range(i + 1, n + 1):
            V = np.zeros((1, 2, n))
            V[0, :, i:j] = rng.uniform(-1, 1, (2, j - i))
            for transpose in (False, True):
                plain = po.decomp_solve(st, data, V.copy(), transpose=transpose)
                ref = po.decomp_solve(st, data, V.copy(), transpose=transpose, start=i, end=j)
                out = g.solveLTranspose(V.copy(), i, j) if transpose else g.solveL(V.copy(), i, j)
                assert np.array_equal(out, ref), (type, i, j, transpose)
                assert np.abs(out - plain).max() <= 1e-12 * max(1.0, np.abs(plain).max())


def test_non_positive_block_is_reported():
    sizes = [12] * 8
    st = Structure.packed(Type.TriBlockDiagonal, sizes)
    H = sc.make_H(Type.TriBlockDiagonal, sizes, 16, seed=1, shift=1.0)
    H[3] -= 100.0 * np.eye(96)
    H[11, 50, 50] = -1.0
    g, ok, data_ref, ok_ref = _factor_both(st, H)
    assert np.array_equal(ok, ok_ref.astype(bool))
    assert (~ok).sum() == 2 and not ok[3] and not ok[11]
    assert np.array_equal(g.data[ok], data_ref[ok_ref.astype(bool)])


def test_dense_layout_keeps_upper_triangle():
    sizes = [4, 6, 3]
    st = Structure.dense(Type.TriBlockDiagonal, sizes)
    H = sc.make_H(Type.TriBlockDiagonal, sizes, 4, seed=8)
    data0 = st.pack(H)
    g = StructuredG(st, data0.copy())
    assert g.lltInPlace().all()
    a, b = g.data.reshape(4, 13, 13), data0.reshape(4, 13, 13)
    iu = np.triu_indices(13, 1)
    assert np.array_equal(a[:, iu[1], iu[0]], b[:, iu[1], iu[0]])  # [col][row] storage: strict upper untouched


@pytest.mark.parametrize("kernel", [1, 2, 3])
@pytest.mark.parametrize("size,blocks,batch", [(12, 32, 1001), (8, 5, 77), (16, 3, 130), (12, 1, 9), (12, 2, 1)])
def test_small_tile_llt_kernels_bit_exact(kernel, size, blocks, batch):
    """The general kernel (1), the small-tile kernel (2: two instances per warp, tiles in registers) and its TMA
    variant (3: next tiles by cp.async.bulk + mbarrier) produce the bits of the oracle — odd batches (a warp with one
    live instance), a single block, and instances that are NOT positive definite (nothing written from the failing
    block on, the other instance of the warp unaffected)."""
    sizes = [size] * blocks
    st = Structure.packed(Type.TriBlockDiagonal, sizes)
    H = sc.make_H(Type.TriBlockDiagonal, sizes, batch, seed=5 + size, shift=1.0)
    bad = sorted({b for b in (0, 3, batch - 1) if b < batch})
    for j, b in enumerate(bad):
        blk = min(blocks - 1, j)  # break a different diagonal block in each of them
        k = blk * size + size // 2
        H[b, k, k] = -abs(H[b, k, k])
    g, ok, data_ref, ok_ref = _factor_both(st, H, kernel)
    assert np.array_equal(ok, ok_ref.astype(bool))
    assert not ok[bad].any() and ok.sum() == batch - len(bad)
    good = ok.astype(bool)
    # (a failed instance holds a partially overwritten block in the reference — Eigen's llt_inplace works in place — and
    # the untouched input here: its contents are unspecified, only the flag is compared)
    assert np.array_equal(g.data[good], data_ref[good]), "factor not bit-identical to the oracle"
    # the solves (general kernel) consume the factor left by any of the three
    rng = np.random.default_rng(0)
    V = rng.uniform(-1, 1, (batch, 1, st.n))
    ref = po.decomp_solve(st, data_ref, V.copy(), transpose=False, nthreads=os.cpu_count())
    assert np.array_equal(g.solveL(V.copy())[good], ref[good])


@pytest.mark.parametrize("kernel", [1, 3])
@pytest.mark.parametrize("size,blocks", [(12, 6), (8, 5), (16, 4), (12, 1)])
def test_small_tile_solve_kernel_hint_windows(kernel, size, blocks):
    """The batch solves of chains of uniform dense tiles (round 2: one column per warp, in place, uniform-pivot recurrence
    with proven quotients, tiles by TMA bulk copies: structured_solve_small_kernel) against the oracle, bit for bit: both
    directions, several columns per instance, and (start, end) hint windows that begin / end in the middle of a tile, on a
    tile boundary and at the ends — beside the general kernel (1) on the same inputs."""
    sizes = [size] * blocks
    n = size * blocks
    st = Structure.packed(Type.TriBlockDiagonal, sizes)
    batch = 37
    H = sc.make_H(Type.TriBlockDiagonal, sizes, batch, seed=11 + size, shift=1.0)
    g, ok, data_ref, ok_ref = _factor_both(st, H, kernel)
    assert ok.all() and np.array_equal(g.data, data_ref)
    rng = np.random.default_rng(4)
    wins = {(0, n), (0, 1), (n - 1, n), (size // 2, n), (0, n - size // 2), (size, n), (0, size)}
    if blocks > 2:
        wins |= {(size + 3, 2 * size + 5), (2 * size, 3 * size), (size - 1, size + 1), (2 * size - 1, n - 1)}
    before = S.launch_count()
    for (i, j) in sorted(wins):
        for ncols in (1, 3):
            V = np.zeros((batch, ncols, n))
            V[:, :, i:j] = rng.uniform(-1, 1, (batch, ncols, j - i))
            for transpose in (False, True):
                ref = po.decomp_solve(st, data_ref, V.copy(), transpose=transpose, start=i, end=j, nthreads=os.cpu_count())
                out = g.solveLTranspose(V.copy(), i, j) if transpose else g.solveL(V.copy(), i, j)
                assert np.array_equal(out, ref), (kernel, size, blocks, i, j, ncols, transpose)
    assert S.launch_count() > before


def test_small_tile_kernel_refused_on_other_structures():
    st = Structure.packed(Type.BlockArrowDown, [12] * 4)
    g = StructuredG(st, st.pack(sc.make_H(Type.BlockArrowDown, [12] * 4, 2, seed=1, shift=1.0)))
    with pytest.raises(RuntimeError):
        g.set_kernel(2)
