"""Host-side mirror of the reference's structured solver over the C-ABI (jrlqp_blockgi_*).

`BlockGISolver` keeps the reference's names and argument meaning (include/jrl-qp/experimental/BlockGISolver.h:24-46,
src/experimental/BlockGISolver.cpp:18-60): solve(G, a, C, bl, bu, xl, xu) with G a structured::StructuredG and C a
structured::StructuredC, then the DualSolver accessors solution(), multipliers(), objectiveValue(), iterations(),
activeSet() (include/jrl-qp/DualSolver.h:26-60). `BatchedBlockGISolver` is the same call over a batch of problems
that share one block structure. Everything runs in libjrlqp_b200.so (jrl-qp_b200/csrc/blockgi.cuh); no CPU fallback.
"""
import ctypes as C

import numpy as np

from . import solver as _solver
from .solver import ActivationStatus, JrlQpError, SolverOptions, TerminationStatus, _Options, _ptr, _Result
from .structured import CStructure, Structure, _CStructure  # noqa: F401


class _CCStructure(C.Structure):
    _fields_ = [("nblocks", C.c_int32), ("nvar", C.c_void_p), ("ncstr", C.c_void_p), ("offset", C.c_void_p), ("ld", C.c_void_p)]


class _BlockProblem(C.Structure):
    _fields_ = [("batch", C.c_int64), ("G", C.c_void_p), ("G_stride", C.c_int64), ("a", C.c_void_p), ("a_stride", C.c_int64),
                ("C", C.c_void_p), ("C_stride", C.c_int64), ("bl", C.c_void_p), ("bl_stride", C.c_int64),
                ("bu", C.c_void_p), ("bu_stride", C.c_int64), ("xl", C.c_void_p), ("xl_stride", C.c_int64),
                ("xu", C.c_void_p), ("xu_stride", C.c_int64)]


class BlockGiInfo(C.Structure):
    _fields_ = [("n", C.c_int32), ("mc", C.c_int32), ("nb", C.c_int32), ("threads", C.c_int32), ("smem_bytes", C.c_int32),
                ("ctas_per_sm", C.c_int32), ("grid", C.c_int32), ("num_sms", C.c_int32),
                ("workspace_bytes_per_cta", C.c_int64), ("g_elements_per_instance", C.c_int64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


EXPORTED_SYMBOLS = ["jrlqp_blockgi_create", "jrlqp_blockgi_destroy", "jrlqp_blockgi_last_error", "jrlqp_blockgi_set_options",
                    "jrlqp_blockgi_get_options", "jrlqp_blockgi_solve_device", "jrlqp_blockgi_solve_host", "jrlqp_blockgi_get_info"]


def _lib():
    lib = _solver.load_library()
    if not getattr(lib, "_blockgi_ready", False):
        lib.jrlqp_blockgi_create.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int32]
        lib.jrlqp_blockgi_destroy.argtypes = [C.c_void_p]
        lib.jrlqp_blockgi_last_error.restype = C.c_char_p
        lib.jrlqp_blockgi_last_error.argtypes = [C.c_void_p]
        lib.jrlqp_blockgi_set_options.argtypes = [C.c_void_p, C.POINTER(_Options)]
        lib.jrlqp_blockgi_get_options.argtypes = [C.c_void_p, C.POINTER(_Options)]
        lib.jrlqp_blockgi_solve_device.argtypes = [C.c_void_p, C.POINTER(_BlockProblem), C.POINTER(_Result), C.c_void_p]
        lib.jrlqp_blockgi_solve_host.argtypes = [C.c_void_p, C.POINTER(_BlockProblem), C.POINTER(_Result)]
        lib.jrlqp_blockgi_get_info.argtypes = [C.c_void_p, C.POINTER(BlockGiInfo)]
        lib._blockgi_ready = True
    return lib


class BatchedBlockGISolver:
    """Batched experimental::BlockGISolver for one block structure. Arrays: Gdata [B, stG.stride] (or [stride]: shared),
    a [B, n], Cdata [B, stC.stride] (or shared), bl/bu [B, mc], xl/xu [B, n] or None; one dimension less = shared."""

    def __init__(self, stG, stC, useBounds, batch_capacity=1, device=0):
        self._lib = _lib()
        self.stG, self.stC = stG, stC
        self.n, self.mc = stG.n, stC.mc
        self.nb = self.n if useBounds else 0
        self.m = self.mc + self.nb
        self.capacity = int(batch_capacity)
        self._keep = [np.ascontiguousarray(stG.sizes, dtype=np.int32), np.ascontiguousarray(stG.diag_offset, dtype=np.int64),
                      np.ascontiguousarray(stG.diag_ld, dtype=np.int32), np.ascontiguousarray(stG.off_offset, dtype=np.int64),
                      np.ascontiguousarray(stG.off_ld, dtype=np.int32), np.ascontiguousarray(stC.nvar, dtype=np.int32),
                      np.ascontiguousarray(stC.ncstr, dtype=np.int32), np.ascontiguousarray(stC.offset, dtype=np.int64),
                      np.ascontiguousarray(stC.ld, dtype=np.int32)]
        k = self._keep
        g = _CStructure(int(stG.type), len(stG.sizes), *[a.ctypes.data for a in k[:5]])
        c = _CCStructure(len(stC.nvar), *[a.ctypes.data for a in k[5:]])
        self._h = C.c_void_p()
        rc = self._lib.jrlqp_blockgi_create(C.byref(self._h), C.byref(g), C.byref(c), int(bool(useBounds)), self.capacity, device)
        if rc != 0:
            msg = self._lib.jrlqp_blockgi_last_error(self._h).decode() if self._h else "allocation failed"
            if self._h:
                self._lib.jrlqp_blockgi_destroy(self._h)
                self._h = None
            raise JrlQpError(f"jrlqp_blockgi_create failed ({rc}): {msg}")
        self._options = SolverOptions()
        self.last = None

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._lib.jrlqp_blockgi_destroy(h)
            self._h = None

    def options(self, opt=None):
        if opt is None:
            return self._options
        self._options = opt
        o = _Options(opt.maxIter_, opt.bigBnd_, int(opt.warmStart_), opt.logFlags_)
        rc = self._lib.jrlqp_blockgi_set_options(self._h, C.byref(o))
        if rc != 0:
            raise JrlQpError(f"jrlqp_blockgi_set_options failed ({rc}): {self._lib.jrlqp_blockgi_last_error(self._h).decode()}")
        return self

    def info(self):
        i = BlockGiInfo()
        self._lib.jrlqp_blockgi_get_info(self._h, C.byref(i))
        return i.as_dict()

    def _problem(self, B, G, a, Cd, bl, bu, xl, xu, shared):
        st = lambda name, per: 0 if name in shared else per  # noqa: E731
        pb = _BlockProblem()
        pb.batch = B
        pb.G, pb.G_stride = _ptr(G), st("G", self.stG.stride)
        pb.a, pb.a_stride = _ptr(a), st("a", self.n)
        pb.C, pb.C_stride = _ptr(Cd), st("C", self.stC.stride)
        pb.bl, pb.bl_stride = _ptr(bl), st("bl", self.mc)
        pb.bu, pb.bu_stride = _ptr(bu), st("bu", self.mc)
        pb.xl, pb.xl_stride = _ptr(xl), st("xl", self.n)
        pb.xu, pb.xu_stride = _ptr(xu), st("xu", self.n)
        return pb

    def solve(self, Gdata, a, Cdata, bl, bu, xl=None, xu=None):
        """HOST arrays. Returns the worst TerminationStatus; results in self.last."""
        f64 = lambda v: None if v is None else np.ascontiguousarray(v, dtype=np.float64)  # noqa: E731
        Gdata, a, Cdata, bl, bu, xl, xu = map(f64, (Gdata, a, Cdata, bl, bu, xl, xu))
        if self.nb == 0:
            xl = xu = None
        arrs = {"G": Gdata, "a": a, "C": Cdata, "bl": bl, "bu": bu, "xl": xl, "xu": xu}
        shared = {k for k, v in arrs.items() if v is not None and v.ndim < 2}
        Bs = [v.shape[0] for k, v in arrs.items() if v is not None and k not in shared]
        B = max(Bs) if Bs else 1
        n, m = self.n, self.m
        x = np.empty((B, n))
        u = np.empty((B, m))
        f = np.empty(B)
        it = np.empty(B, dtype=np.int32)
        status = np.empty(B, dtype=np.int32)
        act = np.empty((B, m), dtype=np.int8)
        alist = np.empty((B, n), dtype=np.int32)
        nact = np.empty(B, dtype=np.int32)
        pb = self._problem(B, Gdata, a, Cdata, bl, bu, xl, xu, shared)
        res = _Result(_ptr(x), _ptr(u), _ptr(f), _ptr(it), _ptr(status), _ptr(act), _ptr(alist), _ptr(nact), None)
        rc = self._lib.jrlqp_blockgi_solve_host(self._h, C.byref(pb), C.byref(res))
        if rc < 0:
            raise JrlQpError(f"jrlqp_blockgi_solve_host failed ({rc}): {self._lib.jrlqp_blockgi_last_error(self._h).decode()}")
        self.last = dict(x=x, u=u, f=f, iterations=it, status=status, active_set=act, active_list=alist, n_active=nact, worst=rc)
        return TerminationStatus(rc)

    def solve_device(self, B, G, a, Cd, bl, bu, xl, xu, x, u=None, f=None, iterations=None, status=None, active_set=None,
                     active_list=None, n_active=None, shared=(), stream=None, G_stride=None):
        """DEVICE pointers (ints, torch tensors); asynchronous on `stream`. `shared`: names of stride-0 arrays.
        G_stride: elements between the instances of G when they are not packed back to back."""
        pb = self._problem(B, G, a, Cd, bl, bu, xl, xu, set(shared))
        if G_stride is not None:
            pb.G_stride = int(G_stride)
        res = _Result(_ptr(x), _ptr(u), _ptr(f), _ptr(iterations), _ptr(status), _ptr(active_set), _ptr(active_list), _ptr(n_active), None)
        rc = self._lib.jrlqp_blockgi_solve_device(self._h, C.byref(pb), C.byref(res), C.c_void_p(stream or 0))
        if rc != 0:
            raise JrlQpError(f"jrlqp_blockgi_solve_device failed ({rc}): {self._lib.jrlqp_blockgi_last_error(self._h).decode()}")


class StructuredC:
    """structured::StructuredC: a CStructure plus the data of ONE instance."""

    def __init__(self, structure, data):
        self.st = structure
        self.data = np.ascontiguousarray(data, dtype=np.float64).ravel()

    def nbVar(self):
        return self.st.n

    def nbCstr(self):
        return self.st.mc


class BlockGISolver:
    """One-problem mirror: solve(G, a, C, bl, bu, xl, xu) with G = (Structure, data) and C a StructuredC."""

    def __init__(self, nbVar=0, nbCstr=0, useBounds=False, device=0):
        self._device = device
        self._solver = None
        self._key = None
        self._opt = SolverOptions()

    def options(self, opt=None):
        if opt is None:
            return self._opt
        self._opt = opt
        if self._solver is not None:
            self._solver.options(opt)
        return self

    def solve(self, G, a, Cs, bl, bu, xl=None, xu=None):
        stG, gdata = G
        xl = None if xl is None or len(xl) == 0 else xl
        xu = None if xu is None or len(xu) == 0 else xu
        key = (id(stG), id(Cs.st), xl is not None)
        if self._key != key:
            self._solver = BatchedBlockGISolver(stG, Cs.st, xl is not None, 1, self._device)
            self._solver.options(self._opt)
            self._key = key
        f1 = lambda v: None if v is None else np.asarray(v, dtype=np.float64)[None, :]  # noqa: E731
        rc = self._solver.solve(f1(gdata), f1(a), f1(Cs.data), f1(bl), f1(bu), f1(xl), f1(xu))
        self._r = self._solver.last
        return rc

    def solution(self):
        return self._r["x"][0]

    def multipliers(self):
        return self._r["u"][0]

    def objectiveValue(self):
        return float(self._r["f"][0])

    def iterations(self):
        return int(self._r["iterations"][0])

    def activeSet(self):
        return [ActivationStatus(int(v)) for v in self._r["active_set"][0]]
