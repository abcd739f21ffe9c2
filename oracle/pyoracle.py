"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

ctypes driver for the CPU oracle (oracle/gi_oracle.cpp). Imported only by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import build as _build  # noqa: E402

_lib = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)
c_bp = C.POINTER(C.c_byte)

STATUS_NAMES = ["SUCCESS", "INCONSISTENT_INPUT", "NON_POS_HESSIAN", "INFEASIBLE", "MAX_ITER_REACHED",
                "LINEAR_DEPENDENCY_DETECTED", "OVERCONSTRAINED_PROBLEM", "UNKNOWN"]
INACTIVE, LOWER, UPPER, EQUALITY, LOWER_BOUND, UPPER_BOUND, FIXED = range(7)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_build.library_path())
        _lib.gi_oracle_dot4.restype = C.c_double
        _lib.gi_oracle_dot32.restype = C.c_double
        _lib.gi_oracle_as_create.restype = C.c_void_p
    return _lib


def _dp(a):
    return None if a is None else a.ctypes.data_as(c_dp)


def _ip(a):
    return None if a is None else a.ctypes.data_as(c_ip)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _stride(a, per):
    """element stride between instances: 0 when the array is shared (ndim one less)."""
    return 0 if a is None or a.size == per else per


_fast = None


def fast_lib():
    """The timing-only build (-O3 -ffast-math, AVX2 + FMA: NOT bit-pinned, see build.py) or None."""
    global _fast
    if _fast is None:
        path = _build.build_fast()
        if path is None:
            return None
        _fast = C.CDLL(path)
    return _fast


def solve_batch(G, a, Cm, bl, bu, xl=None, xu=None, max_iter=500, big_bnd=1e100, nthreads=1,
                want_L=False, instrument=False, experimental=False, warm_start=False, as_in=None, fast=False):
    """G: [B,n,n] (each n x n block column-major, i.e. G[b, j, i] = G_b(i, j); symmetric input makes
    this immaterial), a: [B,n], Cm: [B,mc,n] (row i = constraint normal i = column i of the reference's
    n x mc column-major C), bl/bu: [B,mc], xl/xu: [B,n] or None. Arrays with one dimension less are
    shared by all instances (stride 0). Returns a dict of numpy arrays.
    experimental=True runs the restatement of experimental::GoldfarbIdnaniSolver (warm-start capable):
    as_in [B, mc+nb] int8 is the initial active-set guess (used only with warm_start=True).
    """
    G = _f64(G)
    a = _f64(a)
    Cm = _f64(Cm)
    bl = _f64(bl)
    bu = _f64(bu)
    xl = _f64(xl)
    xu = _f64(xu)
    n = G.shape[-1]
    mc = Cm.shape[-2] if Cm.size else 0
    nb = n if xl is not None and xl.size else 0
    m = mc + nb
    Bs = [arr.shape[0] for arr, nd in ((G, 3), (a, 2), (Cm, 3), (bl, 2), (bu, 2)) if arr.ndim == nd]
    if xl is not None and xl.ndim == 2:
        Bs.append(xl.shape[0])
    B = max(Bs) if Bs else 1
    x = np.empty((B, n))
    u = np.empty((B, m))
    f = np.empty(B)
    iters = np.empty(B, dtype=np.int32)
    status = np.empty(B, dtype=np.int32)
    act = np.empty((B, m), dtype=np.int8)
    alist = np.empty((B, n), dtype=np.int32)
    nact = np.empty(B, dtype=np.int32)
    L = np.empty((B, n, n)) if want_L else None
    flops = np.empty(B) if instrument else None
    margin = np.empty(B) if instrument else None
    if Cm.size == 0:
        Cm = np.zeros((1,))
        bl = np.zeros((1,))
        bu = np.zeros((1,))
    L_ = fast_lib() if fast else lib()  # fast: timing only, results are NOT the canonical bits
    fn = L_.gi_oracle_solve_batch
    pre = []
    if experimental:
        fn = L_.gi_oracle_solve_batch_warm
        if as_in is not None:
            as_in = np.ascontiguousarray(as_in, dtype=np.int8)
            assert as_in.shape[-1] == m
        pre = [None if as_in is None else as_in.ctypes.data_as(c_bp), C.c_long(m if (as_in is not None and as_in.ndim == 2) else 0),
               C.c_int(1 if warm_start else 0)]
    worst = fn(
        *pre, C.c_int(n), C.c_int(mc), C.c_int(nb), C.c_long(B),
        _dp(G), C.c_long(_stride(G, n * n) if G.ndim == 3 else 0), C.c_int(n),
        _dp(a), C.c_long(n if a.ndim == 2 else 0),
        _dp(Cm), C.c_long(mc * n if Cm.ndim == 3 else 0), C.c_int(n),
        _dp(bl), C.c_long(mc if bl.ndim == 2 else 0), _dp(bu), C.c_long(mc if bu.ndim == 2 else 0),
        _dp(xl), C.c_long(n if (xl is not None and xl.ndim == 2) else 0),
        _dp(xu), C.c_long(n if (xu is not None and xu.ndim == 2) else 0),
        C.c_int(max_iter), C.c_double(big_bnd),
        _dp(x), _dp(u), _dp(f), _ip(iters), _ip(status), act.ctypes.data_as(c_bp), _ip(alist), _ip(nact),
        _dp(L), _dp(flops), _dp(margin), C.c_int(nthreads))
    out = dict(x=x, u=u, f=f, iterations=iters, status=status, active_set=act, active_list=alist,
               n_active=nact, worst=worst)
    if want_L:
        out["L"] = L
    if instrument:
        out["flops"] = flops
        out["margin"] = margin
    return out


def solve_trace(G, a, Cm, bl, bu, xl=None, xu=None, max_iter=500, big_bnd=1e100, max_events=4096):
    """Single QP, returns (result dict with 'events' [ne,7] = it,p,status,l,kind,t1,t2 and J, R)."""
    G = _f64(G)
    a = _f64(a)
    Cm = _f64(Cm)
    bl = _f64(bl)
    bu = _f64(bu)
    xl = _f64(xl)
    xu = _f64(xu)
    n = G.shape[-1]
    mc = Cm.shape[0] if Cm.size else 0
    nb = n if xl is not None and xl.size else 0
    m = mc + nb
    x = np.empty(n)
    u = np.empty(m)
    f = C.c_double()
    it = C.c_int()
    ne = C.c_int()
    ev = np.zeros((max_events, 7))
    J = np.empty((n, n))
    R = np.empty((n, n))
    if Cm.size == 0:
        Cm = np.zeros((1,))
        bl = np.zeros((1,))
        bu = np.zeros((1,))
    st = lib().gi_oracle_solve_trace(
        C.c_int(n), C.c_int(mc), C.c_int(nb), _dp(G), C.c_int(n), _dp(a), _dp(Cm), C.c_int(n), _dp(bl), _dp(bu),
        _dp(xl), _dp(xu), C.c_int(max_iter), C.c_double(big_bnd), _dp(x), _dp(u), C.byref(f), C.byref(it),
        _dp(ev), C.c_int(max_events), C.byref(ne), _dp(J), _dp(R))
    return dict(status=st, x=x, u=u, f=f.value, iterations=it.value, events=ev[:min(ne.value, max_events)],
                J=J.T.copy(), R=R.T.copy())


class ActiveSet:
    """Handle on the oracle's ActiveSet restatement (src/internal/ActiveSet.cpp)."""

    def __init__(self, n_cstr, n_bnd=0):
        self.n = n_cstr + n_bnd
        self.h = C.c_void_p(lib().gi_oracle_as_create(C.c_int(n_cstr), C.c_int(n_bnd)))

    def __del__(self):
        if getattr(self, "h", None):
            lib().gi_oracle_as_destroy(self.h)
            self.h = None

    def activate(self, idx, status):
        lib().gi_oracle_as_activate(self.h, C.c_int(idx), C.c_int(status))

    def deactivate(self, active_idx):
        lib().gi_oracle_as_deactivate(self.h, C.c_int(active_idx))

    def reset(self):
        lib().gi_oracle_as_reset(self.h)

    def query(self):
        st = np.zeros(self.n, dtype=np.int8)
        al = np.zeros(max(self.n, 1), dtype=np.int32)
        cnt = np.zeros(8, dtype=np.int32)
        q = lib().gi_oracle_as_query(self.h, st.ctypes.data_as(c_bp), _ip(al), _ip(cnt))
        return st.tolist(), al[:q].tolist(), cnt.tolist()


def dot4(a, b):
    a = _f64(a)
    b = _f64(b)
    return lib().gi_oracle_dot4(C.c_int(a.size), _dp(a), _dp(b))


def dot32(a, b):
    a = _f64(a)
    b = _f64(b)
    return lib().gi_oracle_dot32(C.c_int(a.size), _dp(a), _dp(b))


def givens(p, q):
    out = np.zeros(3)
    lib().gi_oracle_givens(C.c_double(p), C.c_double(q), _dp(out))
    return tuple(out)


# ---------------------------------------------------------------------------------------------
# structured decompositions (oracle/decomp_oracle.cpp)
# ---------------------------------------------------------------------------------------------
c_lp = C.POINTER(C.c_long)


def _desc_args(st):
    """st: any object with type, sizes, diag_offset, diag_ld, off_offset, off_ld (see
    jrl_qp_b200.structured.Structure). Returns (ctypes args, keep-alive list)."""
    size = np.ascontiguousarray(st.sizes, dtype=np.int32)
    doff = np.ascontiguousarray(st.diag_offset, dtype=np.int64)
    dld = np.ascontiguousarray(st.diag_ld, dtype=np.int32)
    ooff = np.ascontiguousarray(st.off_offset, dtype=np.int64)
    old = np.ascontiguousarray(st.off_ld, dtype=np.int32)
    keep = [size, doff, dld, ooff, old]
    args = [C.c_int(int(st.type)), C.c_int(len(size)), _ip(size), doff.ctypes.data_as(c_lp), _ip(dld),
            ooff.ctypes.data_as(c_lp), _ip(old)]
    return args, keep


def decomp_llt(st, data, nthreads=1):
    """In-place structured Cholesky of data [B, stride] (float64, C-contiguous). Returns ok [B] (int32)."""
    assert data.dtype == np.float64 and data.flags.c_contiguous and data.ndim == 2
    B, stride = data.shape
    ok = np.zeros(B, dtype=np.int32)
    args, keep = _desc_args(st)
    lib().decomp_oracle_llt(*args, _dp(data), C.c_long(stride), C.c_long(B), _ip(ok), C.c_int(nthreads))
    return ok


def decomp_solve(st, data, M, transpose=False, start=0, end=-1, nthreads=1):
    """In-place L X = M (or L^T X = M) on M [B, ncols, n] (each instance column-major n x ncols, ld n)."""
    assert data.dtype == np.float64 and data.flags.c_contiguous and data.ndim == 2
    assert M.dtype == np.float64 and M.flags.c_contiguous and M.ndim == 3
    B, stride = data.shape
    ncols, n = M.shape[1], M.shape[2]
    args, keep = _desc_args(st)
    lib().decomp_oracle_solve(*args, _dp(data), C.c_long(stride), _dp(M), C.c_int(n), C.c_int(ncols), C.c_long(ncols * n),
                              C.c_long(B), C.c_int(1 if transpose else 0), C.c_int(start), C.c_int(end), C.c_int(nthreads))
    return M


class CStructure:
    """Block-diagonal constraint matrix (structured::StructuredC, src/structured/StructuredC.cpp:9-25): block i is
    nvar[i] x ncstr[i], column-major (one constraint normal per column) at element offset `offset[i]` from the instance
    base, leading dimension ld[i]. `stride` = elements per instance."""

    def __init__(self, nvar, ncstr, offset, ld, stride):
        self.nvar = np.asarray(nvar, dtype=np.int32)
        self.ncstr = np.asarray(ncstr, dtype=np.int32)
        self.offset = np.asarray(offset, dtype=np.int64)
        self.ld = np.asarray(ld, dtype=np.int32)
        self.stride = int(stride)
        self.n = int(self.nvar.sum())
        self.mc = int(self.ncstr.sum())

    @classmethod
    def packed(cls, nvar, ncstr):
        nvar = np.asarray(nvar, dtype=np.int64)
        ncstr = np.asarray(ncstr, dtype=np.int64)
        off = np.concatenate([[0], np.cumsum(nvar * ncstr)])
        return cls(nvar, ncstr, off[:-1], nvar, int(off[-1]))

    @classmethod
    def dense(cls, nvar, ncstr, ld=None):
        """Blocks are views into a dense column-major n x mc matrix (tests/BlockGISolverTest.in.cpp:90-100)."""
        nvar = np.asarray(nvar, dtype=np.int64)
        ncstr = np.asarray(ncstr, dtype=np.int64)
        n, mc = int(nvar.sum()), int(ncstr.sum())
        ld = n if ld is None else int(ld)
        r0 = np.concatenate([[0], np.cumsum(nvar)])[:-1]
        c0 = np.concatenate([[0], np.cumsum(ncstr)])[:-1]
        return cls(nvar, ncstr, r0 + c0 * ld, np.full(len(nvar), ld), ld * mc)

    def pack(self, Cd):
        """Dense [B, mc, n] (row j = normal of constraint j) -> data [B, stride] in this layout."""
        Cd = np.asarray(Cd, dtype=np.float64)
        B = Cd.shape[0]
        data = np.zeros((B, self.stride))
        r0 = np.concatenate([[0], np.cumsum(self.nvar)])
        c0 = np.concatenate([[0], np.cumsum(self.ncstr)])
        for i in range(len(self.nvar)):
            for j in range(int(self.ncstr[i])):
                o = int(self.offset[i]) + j * int(self.ld[i])
                data[:, o:o + int(self.nvar[i])] = Cd[:, c0[i] + j, r0[i]:r0[i + 1]]
        return data

    def to_dense(self, data):
        """data [B, stride] -> dense [B, mc, n]."""
        B = data.shape[0]
        out = np.zeros((B, self.mc, self.n))
        r0 = np.concatenate([[0], np.cumsum(self.nvar)])
        c0 = np.concatenate([[0], np.cumsum(self.ncstr)])
        for i in range(len(self.nvar)):
            for j in range(int(self.ncstr[i])):
                o = int(self.offset[i]) + j * int(self.ld[i])
                out[:, c0[i] + j, r0[i]:r0[i + 1]] = data[:, o:o + int(self.nvar[i])]
        return out


def block_solve_batch(stG, stC, Gdata, a, Cdata, bl, bu, xl=None, xu=None, max_iter=500, big_bnd=1e100, nthreads=1, want_L=False):
    """Restatement of experimental::BlockGISolver::solve, batched. stG: structure of G (type, sizes, diag_offset, diag_ld,
    off_offset, off_ld, stride), Gdata [B, stride] or [stride] (shared; never modified: the oracle factorises a private copy),
    stC: CStructure, Cdata [B, stride] or [stride], a [B, n] or [n], bl/bu [B, mc] or [mc], xl/xu [B, n], [n] or None."""
    Gdata, a, Cdata, bl, bu, xl, xu = map(_f64, (Gdata, a, Cdata, bl, bu, xl, xu))
    n, mc = stC.n, stC.mc
    assert n == int(np.sum(stG.sizes))
    nb = n if xl is not None else 0
    m = mc + nb
    B = max(arr.shape[0] if arr.ndim == 2 else 1 for arr in (Gdata, a, Cdata, bl, bu))
    out = dict(x=np.zeros((B, n)), u=np.zeros((B, m)), f=np.zeros(B), iterations=np.zeros(B, dtype=np.int32),
               status=np.zeros(B, dtype=np.int32), active_set=np.zeros((B, m), dtype=np.int8),
               active_list=np.zeros((B, n), dtype=np.int32), n_active=np.zeros(B, dtype=np.int32),
               q_doubles=np.zeros(B, dtype=np.int64))
    L = np.zeros((B, stG.stride)) if want_L else None
    args, keep = _desc_args(stG)
    cn, cm = np.ascontiguousarray(stC.nvar, dtype=np.int32), np.ascontiguousarray(stC.ncstr, dtype=np.int32)
    co, cl = np.ascontiguousarray(stC.offset, dtype=np.int64), np.ascontiguousarray(stC.ld, dtype=np.int32)
    st = lambda arr, per: 0 if arr is None or arr.ndim == 1 else per  # noqa: E731
    worst = lib().block_oracle_solve_batch(
        *args, C.c_long(stG.stride), C.c_int(len(cn)), _ip(cn), _ip(cm), co.ctypes.data_as(c_lp), _ip(cl), C.c_int(1 if nb else 0),
        C.c_long(B), _dp(Gdata), C.c_long(st(Gdata, stG.stride)), _dp(a), C.c_long(st(a, n)), _dp(Cdata), C.c_long(st(Cdata, stC.stride)),
        _dp(bl), C.c_long(st(bl, mc)), _dp(bu), C.c_long(st(bu, mc)), _dp(xl), C.c_long(st(xl, n)), _dp(xu), C.c_long(st(xu, n)),
        C.c_int(max_iter), C.c_double(big_bnd), _dp(out["x"]), _dp(out["u"]), _dp(out["f"]), _ip(out["iterations"]), _ip(out["status"]),
        out["active_set"].ctypes.data_as(c_bp), _ip(out["active_list"]), _ip(out["n_active"]), _dp(L),
        out["q_doubles"].ctypes.data_as(c_lp), C.c_int(nthreads))
    out["worst"] = worst
    if want_L:
        out["L"] = L
    return out


def kkt_check_batch(x, u, G, a, Cm, bl, bu, xl=None, xu=None, x_ref=None, tau_p=1e-6, tau_d=1e-6, prec=1e-6, nthreads=1):
    """Restatement of jrl::qp::test::testKKT (src/test/kkt.cpp) for a batch, arrays laid out as in solve_batch.
    Returns (flags [B] int32, resid [B,4], n_fail)."""
    G, a, Cm, bl, bu, xl, xu, x, u, x_ref = map(_f64, (G, a, Cm, bl, bu, xl, xu, x, u, x_ref))
    n = G.shape[-1]
    mc = Cm.shape[-2] if (Cm is not None and Cm.size) else 0
    nb = n if xl is not None and xl.size else 0
    B = x.shape[0]
    flags = np.empty(B, dtype=np.int32)
    resid = np.empty((B, 4))
    if mc == 0:
        Cm = bl = bu = np.zeros((1,))
    nf = lib().kkt_oracle_check_batch(
        C.c_int(n), C.c_int(mc), C.c_int(nb), C.c_long(B),
        _dp(G), C.c_long(n * n if G.ndim == 3 else 0), C.c_int(n),
        _dp(a), C.c_long(n if a.ndim == 2 else 0),
        _dp(Cm), C.c_long(mc * n if Cm.ndim == 3 else 0), C.c_int(n),
        _dp(bl), C.c_long(mc if bl.ndim == 2 else 0), _dp(bu), C.c_long(mc if bu.ndim == 2 else 0),
        _dp(xl), C.c_long(n if (xl is not None and xl.ndim == 2) else 0),
        _dp(xu), C.c_long(n if (xu is not None and xu.ndim == 2) else 0),
        _dp(x), _dp(u), _dp(x_ref), C.c_double(tau_p), C.c_double(tau_d), C.c_double(prec),
        _ip(flags), _dp(resid), C.c_int(nthreads))
    return flags, resid, int(nf)
