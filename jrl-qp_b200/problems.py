"""Synthetic problem batches and KKT verification (host-side test support).

Mirrors the reference's test-support library: `ProblemCharacteristics` / `randomProblem`
(include/jrl-qp/test/randomProblems.h:16-146, src/test/randomProblems.cpp:15-251) and `testKKT`
(src/test/kkt.cpp:14-195, include/jrl-qp/test/kkt.h:83-84). Not on the solve path.
"""
import ctypes as C
import dataclasses
import os

import numpy as np

from . import build as _build

_lib = None
DEFAULT_SEED = 0x6A726C71  # seed = 0x6A726C71 + instance_index


def _tslib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_build.build_testsupport())
    return _lib


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


@dataclasses.dataclass
class ProblemCharacteristics:
    """Same fields and chained-setter names as the reference struct (nObj == rankObj == nVar)."""
    nVar: int
    nEq: int = 0
    nIneq: int = 0
    nStrongActIneq: int = 0
    nWeakActIneq: int = 0
    nStrongActBounds: int = 0
    nWeakActBounds: int = 0
    bounds: bool = False
    doubleSidedIneq: bool = False


@dataclasses.dataclass
class ProblemBatch:
    """Batch of QPs in the layout of the C-ABI (include/jrlqp_b200.h):
    G [B,n,n] (n x n block, column-major == row-major since symmetric), a [B,n],
    C [B,mc,n] (row i = constraint normal i = column i of the reference's n x mc matrix),
    bl, bu [B,mc], xl, xu [B,n] or None, planted x [B,n] and multipliers lam [B,mc+nb]."""
    G: np.ndarray
    a: np.ndarray
    C: np.ndarray
    bl: np.ndarray
    bu: np.ndarray
    xl: np.ndarray
    xu: np.ndarray
    x: np.ndarray = None
    lam: np.ndarray = None

    @property
    def batch(self):
        return self.a.shape[0]

    @property
    def n(self):
        return self.a.shape[1]

    @property
    def mc(self):
        return 0 if self.bl is None else self.bl.shape[-1]  # bl may be shared by the batch (1-D)

    @property
    def nb(self):
        return 0 if self.xl is None else self.n

    def slice(self, lo, hi):
        s = lambda v: None if v is None else v[lo:hi]
        return ProblemBatch(*(s(getattr(self, f.name)) for f in dataclasses.fields(self)))

    def input_bytes(self):
        tot = 0
        for name in ("G", "a", "C", "bl", "bu", "xl", "xu"):
            v = getattr(self, name)
            tot += 0 if v is None else v.nbytes
        return tot


def random_problems(ch, batch, seed=DEFAULT_SEED, first_index=0, nthreads=None, out=None):
    """Generate `batch` seeded random problems (instance k <- stream seed + first_index + k)."""
    n, mc = ch.nVar, ch.nEq + ch.nIneq
    nb = n if ch.bounds else 0
    if nthreads is None:
        nthreads = min(os.cpu_count() or 1, 64, max(1, batch))
    alloc = (lambda *s: np.empty(s)) if out is None else out
    G = alloc(batch, n, n)
    a = alloc(batch, n)
    Cm = alloc(batch, mc, n)
    bl = alloc(batch, mc)
    bu = alloc(batch, mc)
    xl = alloc(batch, n) if nb else None
    xu = alloc(batch, n) if nb else None
    x = np.empty((batch, n))
    lam = np.empty((batch, mc + nb))
    rc = _tslib().jrlqp_ts_random_problems(
        C.c_int(n), C.c_int(ch.nEq), C.c_int(ch.nIneq), C.c_int(ch.nStrongActIneq), C.c_int(ch.nWeakActIneq),
        C.c_int(ch.nStrongActBounds), C.c_int(ch.nWeakActBounds), C.c_int(int(ch.bounds)),
        C.c_int(int(ch.doubleSidedIneq)), C.c_ulonglong(seed), C.c_long(first_index), C.c_long(batch),
        _dp(G), _dp(a), _dp(Cm), _dp(bl), _dp(bu), _dp(xl), _dp(xu), _dp(x), _dp(lam), C.c_int(nthreads))
    if rc != 0:
        raise ValueError("inconsistent ProblemCharacteristics")
    return ProblemBatch(G, a, Cm, bl, bu, xl, xu, x, lam)


# Named shapes of BASELINE.json / SURVEY.md §8(d). Active fractions follow benchmarks/Solvers.cpp:621-623
# (30 % of min(n, nIneq) inequalities, 10 % of the bounds, double-sided inequalities).
def config_A():  # n=50, 20 eq + 30 ineq + 50 bounds (m=100)  -- the headline metric
    return ProblemCharacteristics(50, 20, 30, 9, 0, 5, 0, True, True)


def config_B():  # n=20, 4 eq + 16 ineq + 20 bounds (m=40)
    return ProblemCharacteristics(20, 4, 16, 5, 0, 2, 0, True, True)


def config_D():  # n=128, 26 eq + 102 ineq + 128 bounds (m=256)
    return ProblemCharacteristics(128, 26, 102, 30, 0, 13, 0, True, True)


def _kkt_constraint(cx, bl, bu, u, tau_x, tau_u):
    # src/test/kkt.cpp:14-22
    li = cx - bl
    ui = cx - bu
    b1 = (np.abs(li) <= tau_x) & (u <= -tau_u)
    b2 = (li >= -tau_x) & (ui <= tau_x) & (np.abs(u) <= tau_u)
    b3 = (np.abs(ui) <= tau_x) & (u >= tau_u)
    return b1 | b2 | b3


def test_kkt(x, u, pb, tau_p=1e-6, tau_d=1e-6):
    """Vectorised testKKT over a batch (src/test/kkt.cpp:84-195): returns a bool array [B].
    x [B,n], u [B,mc+nb] with the reference's sign convention."""
    n, mc = pb.n, pb.mc
    G = pb.G if pb.G.ndim == 3 else pb.G[None]
    Cm = pb.C if pb.C.ndim == 3 else pb.C[None]
    tau_x = tau_p * (1 + np.abs(x).max(axis=1))
    tau_u = tau_d * (1 + (np.abs(u).max(axis=1) if u.shape[1] else 0.0))
    dL = np.einsum("bij,bj->bi", G, x) + pb.a
    if pb.xl is not None:
        dL = dL + u[:, mc:]
    if mc:
        dL = dL + np.einsum("bin,bi->bn", Cm, u[:, :mc])
    ok = np.abs(dL).max(axis=1) <= tau_u
    if mc:
        cx = np.einsum("bin,bn->bi", Cm, x)
        ok &= _kkt_constraint(cx, pb.bl, pb.bu, u[:, :mc], tau_x[:, None], tau_u[:, None]).all(axis=1)
    if pb.xl is not None:
        ok &= _kkt_constraint(x, pb.xl, pb.xu, u[:, mc:], tau_x[:, None], tau_u[:, None]).all(axis=1)
    return ok


def is_approx(a, b, prec=1e-6):
    """Eigen isApprox per row: ||a-b|| <= prec * min(||a||, ||b||)."""
    na = np.linalg.norm(a, axis=-1)
    nbn = np.linalg.norm(b, axis=-1)
    return np.linalg.norm(a - b, axis=-1) <= prec * np.minimum(na, nbn)
