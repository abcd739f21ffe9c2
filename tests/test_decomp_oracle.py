"""CPU tests of the structured-decomposition oracle against the reference's own acceptance criteria:
tests/triBlockDiagLLTTest.cpp:35-87 and tests/blockArrowLLTTest.cpp:39-176 — the structured factor
equals the dense Cholesky factor and the structured solves equal dense triangular solves to 1e-8,
for every (start, end) window of non-zero rows."""
import os
import sys

import numpy as np
import pytest
import scipy.linalg as sla

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import pyoracle as po  # noqa: E402
import structured_cases as sc  # noqa: E402
from jrl_qp_b200.structured import Structure, Type  # noqa: E402

SIZES = [3, 5, 2, 3]  # the block sizes of the reference tests


def is_approx(a, b, prec=1e-8):
    # Eigen isApprox: ||a - b||_F <= prec * min(||a||_F, ||b||_F)
    return np.linalg.norm(a - b) <= prec * min(np.linalg.norm(a), np.linalg.norm(b))


@pytest.mark.parametrize("layout", ["dense", "packed"])
@pytest.mark.parametrize("type", [Type.TriBlockDiagonal, Type.BlockArrowDown, Type.BlockArrowUp])
def test_llt_matches_dense(type, layout):
    st = Structure.dense(type, SIZES) if layout == "dense" else Structure.packed(type, SIZES)
    H = sc.make_H(type, SIZES, 16, seed=7)
    data = st.pack(H)
    ok = po.decomp_llt(st, data)
    assert ok.all()
    L = sc.factor_from_data(st, data)
    Lref = sc.dense_factor(type, SIZES, H)
    for k in range(H.shape[0]):
        assert is_approx(L[k], Lref[k])


def test_llt_dense_layout_leaves_upper_untouched():
    st = Structure.dense(Type.TriBlockDiagonal, SIZES)
    H = sc.make_H(Type.TriBlockDiagonal, SIZES, 2, seed=3)
    data = st.pack(H)
    before = data.copy().reshape(2, 13, 13)  # [b][col][row]
    po.decomp_llt(st, data)
    after = data.reshape(2, 13, 13)
    for c in range(13):
        for r in range(c):
            assert (after[:, c, r] == before[:, c, r]).all()  # "upper part remains whatever was there"


def test_llt_reports_non_positive_block():
    st = Structure.packed(Type.TriBlockDiagonal, SIZES)
    H = sc.make_H(Type.TriBlockDiagonal, SIZES, 3, seed=5)
    H[1] -= 50.0 * np.eye(13)
    ok = po.decomp_llt(st, st.pack(H))
    assert ok.tolist() == [1, 0, 1]


@pytest.mark.parametrize("type", [Type.TriBlockDiagonal, Type.BlockArrowDown, Type.BlockArrowUp])
def test_solves_match_dense_for_every_window(type):
    st = Structure.packed(type, SIZES)
    H = sc.make_H(type, SIZES, 1, seed=11)
    data = st.pack(H)
    assert po.decomp_llt(st, data).all()
    Lref = sc.dense_factor(type, SIZES, H)[0]
    P = sc.permutation_up(SIZES) if type == Type.BlockArrowUp else np.eye(13)
    rng = np.random.default_rng(5)
    n = 13
    for i in range(n):
        for j in range(i + 1, n + 1):
            Bm = np.zeros((n, 5))
            Bm[i:j] = rng.uniform(-1, 1, (j - i, 5))
            # L solve: P L X = B
            X0 = sla.solve_triangular(Lref, P.T @ Bm, lower=True)
            for hints in ((0, -1), (i, j)):
                M = np.ascontiguousarray(Bm.T[None].copy())
                po.decomp_solve(st, data, M, transpose=False, start=hints[0], end=hints[1])
                assert is_approx(M[0].T, X0), (type, i, j, hints)
            # L^T solve: L^T P^T X = B
            X1 = P @ sla.solve_triangular(Lref.T, Bm, lower=False)
            for hints in ((0, -1), (i, j)):
                M = np.ascontiguousarray(Bm.T[None].copy())
                po.decomp_solve(st, data, M, transpose=True, start=hints[0], end=hints[1])
                assert is_approx(M[0].T, X1), (type, i, j, hints)
            # both: H^-1 B  (tests/blockArrowLLTTest.cpp:162-171)
            M = np.ascontiguousarray(Bm.T[None].copy())
            po.decomp_solve(st, data, M, transpose=False)
            po.decomp_solve(st, data, M, transpose=True)
            assert is_approx(M[0].T, np.linalg.solve(H[0], Bm), 1e-7)


@pytest.mark.parametrize("type,sizes", [(Type.TriBlockDiagonal, [12] * 32), (Type.BlockArrowDown, [12] * 8),
                                        (Type.BlockArrowUp, [42] * 5), (Type.TriBlockDiagonal, [43] * 9),
                                        (Type.BlockArrowUp, [4, 7, 3]), (Type.BlockArrowDown, [1, 1])])
def test_llt_other_shapes(type, sizes):
    st = Structure.packed(type, sizes)
    H = sc.make_H(type, sizes, 4, seed=13, shift=1.0)
    data = st.pack(H)
    assert po.decomp_llt(st, data, nthreads=2).all()
    L = sc.factor_from_data(st, data)
    Lref = sc.dense_factor(type, sizes, H)
    assert np.abs(L - Lref).max() <= 1e-8 * np.abs(Lref).max()
    v = np.random.default_rng(1).uniform(-1, 1, (4, st.n))
    M = v[:, None, :].copy()
    po.decomp_solve(st, data, M, transpose=False)
    po.decomp_solve(st, data, M, transpose=True)
    ref = np.linalg.solve(H, v[:, :, None])[:, :, 0]
    assert np.abs(M[:, 0, :] - ref).max() <= 1e-8 * max(1.0, np.abs(ref).max())
