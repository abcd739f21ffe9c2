"""Host-side mirror of the reference's structured Cholesky interface over the C-ABI.

`StructuredG` keeps the reference's names and argument meaning
(include/jrl-qp/structured/StructuredG.h:14-76, src/structured/StructuredG.cpp): a type tag
(TriBlockDiagonal / BlockArrowUp / BlockArrowDown), lltInPlace(), solveInPlaceLTranspose(v),
solveL(out, in) and the solveL overload for a vector with a single non-zero segment — batched: one
object holds `batch` matrices of the same block structure. The free functions mirror
include/jrl-qp/decomposition/{triBlockDiagLLT,blockArrowLLT}.h. Everything runs in
libjrlqp_b200.so (jrl-qp_b200/csrc/structured.cu); there is no CPU fallback.
"""
import ctypes as C
import enum

import numpy as np

from . import solver as _solver


class Type(enum.IntEnum):  # structured::StructuredG::Type, include/jrl-qp/structured/StructuredG.h:17-22
    TriBlockDiagonal = 0
    BlockArrowUp = 1
    BlockArrowDown = 2


class Structure:
    """Where the blocks of one instance live: for every diagonal block D_i (n_i x n_i) and every
    off-diagonal block an element offset from the instance base and a leading dimension (the blocks
    are column-major views, as the reference's std::vector<MatrixRef>). Off-diagonal block i is
      TriBlockDiagonal: S_i, n_{i+1} x n_i      (block (i+1, i))
      BlockArrowDown:   S_i, n_{b-1} x n_i      (last block row)
      BlockArrowUp:     S_i, n_{i+1} x n_0      (first block column)
    """

    def __init__(self, type, sizes, diag_offset, diag_ld, off_offset, off_ld, stride):
        self.type = Type(type)
        self.sizes = np.asarray(sizes, dtype=np.int32)
        self.diag_offset = np.asarray(diag_offset, dtype=np.int64)
        self.diag_ld = np.asarray(diag_ld, dtype=np.int32)
        self.off_offset = np.asarray(off_offset, dtype=np.int64)
        self.off_ld = np.asarray(off_ld, dtype=np.int32)
        self.stride = int(stride)
        self.n = int(self.sizes.sum())
        self.starts = np.concatenate([[0], np.cumsum(self.sizes)]).astype(np.int64)

    def off_shape(self, i):
        s = self.sizes
        if self.type == Type.TriBlockDiagonal:
            return int(s[i + 1]), int(s[i])
        if self.type == Type.BlockArrowDown:
            return int(s[-1]), int(s[i])
        return int(s[i + 1]), int(s[0])

    def off_position(self, i):
        """(row, col) of off-diagonal block i in the full n x n matrix."""
        st = self.starts
        if self.type == Type.TriBlockDiagonal:
            return int(st[i + 1]), int(st[i])
        if self.type == Type.BlockArrowDown:
            return int(st[-2]), int(st[i])
        return int(st[i + 1]), 0

    @classmethod
    def dense(cls, type, sizes, ld=None):
        """Blocks are views into a dense column-major n x n matrix (as the reference's tests pack
        H.block(...) views, tests/triBlockDiagLLTTest.cpp:41-43); stride = ld * n."""
        sizes = np.asarray(sizes, dtype=np.int64)
        n = int(sizes.sum())
        ld = n if ld is None else int(ld)
        self = cls(type, sizes, np.zeros(len(sizes)), np.full(len(sizes), ld), np.zeros(len(sizes) - 1),
                   np.full(len(sizes) - 1, ld), ld * n)
        st = self.starts
        self.diag_offset = (st[:-1] + st[:-1] * ld).astype(np.int64)
        pos = [self.off_position(i) for i in range(len(sizes) - 1)]
        self.off_offset = np.array([r + c * ld for r, c in pos], dtype=np.int64)
        return self

    @classmethod
    def packed(cls, type, sizes):
        """Blocks stored back to back (diagonal blocks first, then the off-diagonal ones), each
        column-major with ld = its row count: contiguous tiles, the layout the CUDA kernels like best."""
        sizes = np.asarray(sizes, dtype=np.int64)
        b = len(sizes)
        self = cls(type, sizes, np.zeros(b), sizes, np.zeros(b - 1), np.ones(b - 1), 0)
        o = 0
        doff = []
        for s in sizes:
            doff.append(o)
            o += int(s) * int(s)
        ooff, old = [], []
        for i in range(b - 1):
            r, c = self.off_shape(i)
            ooff.append(o)
            old.append(r)
            o += r * c
        self.diag_offset = np.array(doff, dtype=np.int64)
        self.off_offset = np.array(ooff, dtype=np.int64)
        self.off_ld = np.array(old, dtype=np.int32)
        self.stride = o
        return self

    # -- conversions used by tests / examples
    def pack(self, H):
        """Dense symmetric H [B, n, n] -> data [B, stride] in this structure's layout."""
        H = np.asarray(H, dtype=np.float64)
        B = H.shape[0]
        data = np.zeros((B, self.stride))
        st = self.starts
        for i, s in enumerate(self.sizes):
            blk = H[:, st[i]:st[i + 1], st[i]:st[i + 1]]
            self._put(data, self.diag_offset[i], self.diag_ld[i], blk)
        for i in range(len(self.sizes) - 1):
            r, c = self.off_position(i)
            nr, nc = self.off_shape(i)
            self._put(data, self.off_offset[i], self.off_ld[i], H[:, r:r + nr, c:c + nc])
        return data

    def unpack_lower(self, data):
        """data [B, stride] -> dense [B, n, n] holding the lower triangles of the diagonal blocks
        and the off-diagonal blocks at their place below the diagonal (zeros elsewhere)."""
        B = data.shape[0]
        out = np.zeros((B, self.n, self.n))
        st = self.starts
        for i, s in enumerate(self.sizes):
            blk = self._get(data, self.diag_offset[i], self.diag_ld[i], int(s), int(s))
            out[:, st[i]:st[i + 1], st[i]:st[i + 1]] = np.tril(blk)
        for i in range(len(self.sizes) - 1):
            r, c = self.off_position(i)
            nr, nc = self.off_shape(i)
            out[:, r:r + nr, c:c + nc] = self._get(data, self.off_offset[i], self.off_ld[i], nr, nc)
        return out

    @staticmethod
    def _put(data, off, ld, blk):
        nr, nc = blk.shape[1], blk.shape[2]
        for c in range(nc):
            data[:, off + c * ld: off + c * ld + nr] = blk[:, :, c]

    @staticmethod
    def _get(data, off, ld, nr, nc):
        out = np.empty((data.shape[0], nr, nc))
        for c in range(nc):
            out[:, :, c] = data[:, off + c * ld: off + c * ld + nr]
        return out


class _CStructure(C.Structure):
    _fields_ = [("type", C.c_int32), ("nblocks", C.c_int32), ("block_size", C.c_void_p),
                ("diag_offset", C.c_void_p), ("diag_ld", C.c_void_p), ("off_offset", C.c_void_p), ("off_ld", C.c_void_p)]


class StructuredG:
    """Batched structured::StructuredG. `data` [batch, stride] float64 holds the matrices in the
    layout described by `structure`; lltInPlace() factorises all of them in place on the GPU."""

    def __init__(self, structure, data, device=0):
        self.st = structure
        self.data = np.ascontiguousarray(data, dtype=np.float64)
        if self.data.ndim == 1:
            self.data = self.data[None, :]
        assert self.data.shape[1] == structure.stride
        self.batch = self.data.shape[0]
        self._lib = _solver.load_library()
        self._keep = [np.ascontiguousarray(structure.sizes, dtype=np.int32),
                      np.ascontiguousarray(structure.diag_offset, dtype=np.int64),
                      np.ascontiguousarray(structure.diag_ld, dtype=np.int32),
                      np.ascontiguousarray(structure.off_offset, dtype=np.int64),
                      np.ascontiguousarray(structure.off_ld, dtype=np.int32)]
        cs = _CStructure(int(structure.type), len(structure.sizes), *[a.ctypes.data for a in self._keep])
        self._h = C.c_void_p()
        rc = self._lib.jrlqp_structured_create(C.byref(self._h), C.byref(cs), C.c_int64(self.batch), C.c_int32(device))
        if rc != 0:
            msg = self._lib.jrlqp_structured_last_error(self._h).decode() if self._h else ""
            if self._h:
                self._lib.jrlqp_structured_destroy(self._h)
                self._h = None
            raise RuntimeError(f"jrlqp_structured_create failed ({rc}): {msg}")
        self.decomposed_ = np.zeros(self.batch, dtype=np.int32)

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.jrlqp_structured_destroy(self._h)
            self._h = None

    def type(self):
        return self.st.type

    def nbVar(self, i=None):
        return self.st.n if i is None else int(self.st.sizes[i])

    def set_kernel(self, mode):
        """jrlqp_structured_set_kernel: 0 automatic, 1 general kernel, 2 small-tile kernel, 3 small-tile kernel + TMA."""
        self._lib.jrlqp_structured_set_kernel.argtypes = [C.c_void_p, C.c_int32]
        if self._lib.jrlqp_structured_set_kernel(self._h, int(mode)) != 0:
            raise RuntimeError(self._lib.jrlqp_structured_last_error(self._h).decode())
        return self

    def lltInPlace(self):
        """bool per instance (the reference returns false when a diagonal block is not positive definite)."""
        rc = self._lib.jrlqp_structured_llt_host(self._h, self.data.ctypes.data_as(C.c_void_p), C.c_int64(self.st.stride),
                                                 C.c_int64(self.batch), self.decomposed_.ctypes.data_as(C.c_void_p))
        if rc < 0:
            raise RuntimeError(f"jrlqp_structured_llt_host failed ({rc}): {self._lib.jrlqp_structured_last_error(self._h).decode()}")
        return self.decomposed_.astype(bool)

    def decomposed(self):
        return self.decomposed_.astype(bool)

    def _solve(self, M, transpose, start, end):
        M = np.ascontiguousarray(M, dtype=np.float64)
        squeeze = M.ndim == 2
        if squeeze:
            M = M[:, None, :]
        assert M.shape[0] == self.batch and M.shape[2] == self.st.n
        rc = self._lib.jrlqp_structured_solve_host(self._h, self.data.ctypes.data_as(C.c_void_p), C.c_int64(self.st.stride),
                                                   M.ctypes.data_as(C.c_void_p), C.c_int32(self.st.n), C.c_int32(M.shape[1]),
                                                   C.c_int64(M.shape[1] * self.st.n), C.c_int64(self.batch),
                                                   C.c_int32(1 if transpose else 0), C.c_int32(start), C.c_int32(end))
        if rc < 0:
            raise RuntimeError(f"jrlqp_structured_solve_host failed ({rc}): {self._lib.jrlqp_structured_last_error(self._h).decode()}")
        return M[:, 0, :] if squeeze else M

    def solveInPlaceLTranspose(self, v):
        """v [batch, n] (or [batch, ncols, n], column-major instances): returns L^-T P^T v."""
        return self._solve(v, True, 0, -1)

    def solveL(self, v, start=0, end=-1):
        """out = (P L)^-1 v; start/end = the SingleNZSegmentVector hints of the reference's overload."""
        return self._solve(v, False, start, end)

    def solveLTranspose(self, v, start=0, end=-1):
        return self._solve(v, True, start, end)


class CStructure:
    """Where the blocks of a block-diagonal constraint matrix live (structured::StructuredC built from a
    std::vector<MatrixConstRef>, src/structured/StructuredC.cpp:9-25): block i is nvar[i] x ncstr[i], column-major
    (one constraint normal per column) at element `offset[i]` from the instance base, leading dimension ld[i];
    `stride` = elements per instance. Constraints are numbered block after block."""

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
        """Blocks stored back to back, ld = rows."""
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
        data = np.zeros((Cd.shape[0], self.stride))
        r0 = np.concatenate([[0], np.cumsum(self.nvar)])
        c0 = np.concatenate([[0], np.cumsum(self.ncstr)])
        for i in range(len(self.nvar)):
            for j in range(int(self.ncstr[i])):
                o = int(self.offset[i]) + j * int(self.ld[i])
                data[:, o:o + int(self.nvar[i])] = Cd[:, c0[i] + j, r0[i]:r0[i + 1]]
        return data
