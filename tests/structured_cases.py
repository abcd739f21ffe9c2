"""Shared generators for the structured-decomposition tests: restatements of the matrix builders of
the reference's tests (tests/triBlockDiagLLTTest.cpp:16-33 biBlockDiagRandom,
tests/blockArrowLLTTest.cpp:17-37 blockDiagAndOneColDiagRandom), seeded."""
import numpy as np

import jrl_qp_b200  # noqa: F401
from jrl_qp_b200.structured import Structure, Type


def bi_block_diag_random(rng, sizes, batch):
    s = int(np.sum(sizes))
    A = np.zeros((batch, s, s))
    k = 0
    for i in range(len(sizes) - 1):
        ni = sizes[i]
        A[:, k:k + ni, k:k + ni] = rng.uniform(-1, 1, (batch, ni, ni))
        A[:, k + ni:k + ni + sizes[i + 1], k:k + ni] = rng.uniform(-1, 1, (batch, sizes[i + 1], ni))
        k += ni
    nb = sizes[-1]
    A[:, s - nb:, s - nb:] = rng.uniform(-1, 1, (batch, nb, nb))
    return A


def block_diag_and_one_col_random(rng, sizes, first, batch):
    s = int(np.sum(sizes))
    A = np.zeros((batch, s, s))
    k = 0
    for ni in sizes:
        A[:, k:k + ni, k:k + ni] = rng.uniform(-1, 1, (batch, ni, ni))
        if first:
            A[:, k:k + ni, 0:sizes[0]] = rng.uniform(-1, 1, (batch, ni, sizes[0]))
        else:
            A[:, k:k + ni, s - sizes[-1]:] = rng.uniform(-1, 1, (batch, ni, sizes[-1]))
        k += ni
    nb = sizes[-1]
    A[:, s - nb:, s - nb:] = rng.uniform(-1, 1, (batch, nb, nb))
    return A


def make_H(type, sizes, batch, seed, shift=0.0):
    """Dense SPD matrices [batch, n, n] with the sparsity of `type` (H = A A^T or A^T A as in the
    reference tests; `shift` adds a multiple of the identity to keep large random cases well conditioned)."""
    rng = np.random.default_rng(seed)
    sizes = list(sizes)
    if type == Type.TriBlockDiagonal:
        A = bi_block_diag_random(rng, sizes, batch)
        H = A @ A.transpose(0, 2, 1)
    else:
        A = block_diag_and_one_col_random(rng, sizes, type == Type.BlockArrowUp, batch)
        H = A.transpose(0, 2, 1) @ A
    n = H.shape[1]
    return H + shift * np.eye(n)[None]


def permutation_up(sizes):
    """P of include/jrl-qp/decomposition/blockArrowLLT.h:36-47: P^T H P moves block 0 last."""
    n = int(np.sum(sizes))
    n0 = sizes[0]
    P = np.zeros((n, n))
    P[:n0, n - n0:] = np.eye(n0)
    P[n0:, :n - n0] = np.eye(n - n0)
    return P


def dense_factor(type, sizes, H):
    """Reference answer: dense Cholesky of H (of P^T H P for the up arrow), [batch, n, n]."""
    if type == Type.BlockArrowUp:
        P = permutation_up(sizes)
        H = P.T[None] @ H @ P[None]
    return np.linalg.cholesky(H)


def factor_from_data(st, data):
    """The full lower-triangular factor [batch, n, n] that `data` (after lltInPlace) represents."""
    sizes = [int(s) for s in st.sizes]
    low = st.unpack_lower(data)
    if st.type != Type.BlockArrowUp:
        return low
    # up: L = [diag(L_2..L_b) 0; B_1..B_{b-1} L_1] with side[i] = B_i^T stored at (block i+1, block 0)
    n = st.n
    n0 = sizes[0]
    B = data.shape[0]
    L = np.zeros((B, n, n))
    L[:, :n - n0, :n - n0] = low[:, n0:, n0:]
    L[:, n - n0:, n - n0:] = low[:, :n0, :n0]
    L[:, n - n0:, :n - n0] = low[:, n0:, :n0].transpose(0, 2, 1)
    return L
