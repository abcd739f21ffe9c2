// Structured Cholesky decompositions and their triangular solves, batched: ONE INSTANCE PER CTA.
//
// Replaces, for a batch of matrices sharing one block structure, the reference functions
//   decomposition::triBlockDiagLLT / triBlockDiagLSolve / triBlockDiagLTransposeSolve
//       (src/decomposition/triBlockDiagLLT.cpp:9-158)
//   decomposition::blockArrowLLT / blockArrowLSolve / blockArrowLTransposeSolve, up and down
//       (src/decomposition/blockArrowLLT.cpp:52-277)
// behind structured::StructuredG (src/structured/StructuredG.cpp:22-113).
//
// Shape of the work: per instance a chain of small dense tiles (config E: 32 diagonal + 31
// sub-diagonal 12 x 12 blocks = 72.6 KB) that is read once, updated and written once — about one flop
// per byte, i.e. HBM-bound. The chain is serial inside an instance (block i+1 needs block i), so
// the parallelism comes from the batch: a CTA is one warp (two / four for blocks wider than 32 / 64
// rows), up to 32 CTAs are resident per SM, and the global-memory latency of one instance's tile
// loads hides behind the arithmetic of the others. Tiles are staged in shared memory column-major
// (lanes run down a column: conflict-free, coalesced against the column-major blocks in HBM).
//
// Arithmetic: the canonical per-output order documented in oracle/decomp_oracle.hpp (dot4 for every
// inner product, column-oriented substitution with true division for the vector solves), so the
// results are bit-identical to the CPU oracle whatever the thread mapping.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace jrlqp
{

enum : int
{
  SG_TRI = 0, // structured::StructuredG::Type::TriBlockDiagonal
  SG_ARROW_UP = 1,
  SG_ARROW_DOWN = 2
};

struct StructParams
{
  int type, b, n, nmax;
  const int * size; // [b]      device
  const long long * doff; // [b]
  const int * dld; // [b]
  const long long * ooff; // [b-1]
  const int * old; // [b-1]
  const int * start; // [b+1] first row of every block
  double * data;
  long long stride;
  long long batch;
  int * ok;
  // solves
  double * M;
  int ldm, ncols;
  long long mstride;
  int transpose, hint_start, hint_end;
};

// strided view of a tile in shared memory: element (r, c) at p[r * rs + c * cs]
struct TView
{
  double * p;
  int rs, cs;
  __device__ __forceinline__ double & operator()(int r, int c) const { return p[r * rs + c * cs]; }
  __device__ __forceinline__ TView t() const { return {p, cs, rs}; }
  __device__ __forceinline__ TView sub(int r0, int c0) const { return {p + r0 * rs + c0 * cs, rs, cs}; }
};

// dot4 over k < len of A(ra, k) * B(rb, k) (two rows of strided tiles)
__device__ __forceinline__ double dot4_rows(int len, const double * a, int sa, const double * b, int sb)
{
  double c0 = 0, c1 = 0, c2 = 0, c3 = 0;
  int k = 0;
  for(; k + 3 < len; k += 4)
  {
    c0 = fma(a[k * sa], b[k * sb], c0);
    c1 = fma(a[(k + 1) * sa], b[(k + 1) * sb], c1);
    c2 = fma(a[(k + 2) * sa], b[(k + 2) * sb], c2);
    c3 = fma(a[(k + 3) * sa], b[(k + 3) * sb], c3);
  }
  if(k < len) c0 = fma(a[k * sa], b[k * sb], c0);
  if(k + 1 < len) c1 = fma(a[(k + 1) * sa], b[(k + 1) * sb], c1);
  if(k + 2 < len) c2 = fma(a[(k + 2) * sa], b[(k + 2) * sb], c2);
  return (c0 + c1) + (c2 + c3);
}

// global (column-major, ld) -> shared tile (column-major, ld = rows), optionally transposed on the way
__device__ __forceinline__ void load_tile(double * dst, const double * __restrict__ src, int rows, int cols, int ld, bool transpose)
{
  const int total = rows * cols;
  for(int idx = threadIdx.x; idx < total; idx += blockDim.x)
  {
    const int c = idx / rows, r = idx - c * rows;
    dst[transpose ? c + r * cols : idx] = src[r + (long long)c * ld];
  }
}

__device__ __forceinline__ void store_tile(double * __restrict__ dst, const double * src, int rows, int cols, int ld, bool transpose, bool lower_only)
{
  const int total = rows * cols;
  for(int idx = threadIdx.x; idx < total; idx += blockDim.x)
  {
    const int c = idx / rows, r = idx - c * rows;
    if(lower_only && r < c) continue; // "its upper part remains whatever was there originally"
    dst[r + (long long)c * ld] = src[transpose ? c + r * cols : idx];
  }
}

// Eigen llt_inplace restated: left-looking, thread = row. D: n x n tile, ld n; vd: n doubles of
// scratch (the pivots v_k, so that the diagonal can be overwritten while others still need v_k).
// Uniform return.
__device__ __forceinline__ bool tile_chol(double * D, int n, double * vd)
{
  for(int k = 0; k < n; ++k)
  {
    for(int i = k + threadIdx.x; i < n; i += blockDim.x)
    {
      const double v = D[i + k * n] - dot4_rows(k, D + i, n, D + k, n);
      D[i + k * n] = v;
      if(i == k) vd[k] = v;
    }
    __syncthreads();
    const double vk = vd[k];
    if(vk <= 0.0) return false; // Eigen: "if (x <= 0) return k"
    const double lkk = sqrt(vk);
    for(int i = k + threadIdx.x; i < n; i += blockDim.x) D[i + k * n] = i == k ? lkk : D[i + k * n] / lkk;
    __syncthreads();
  }
  return true;
}

// B = B L^-T, thread = row of B (B: rows x n tile, ld rows; L: n x n tile, ld n). No barrier needed
// inside: a row only depends on itself and on L.
__device__ __forceinline__ void tile_trsm_right_lt(double * B, int rows, const double * L, int n)
{
  for(int r = threadIdx.x; r < rows; r += blockDim.x)
    for(int k = 0; k < n; ++k) B[r + k * rows] = (B[r + k * rows] - dot4_rows(k, B + r, rows, L + k, n)) / L[k + k * n];
}

// D -= B B^T on the lower triangle (D: n x n, ld n; B: n x kk, ld n), outputs spread over the threads
__device__ __forceinline__ void tile_syrk_sub(double * D, int n, const double * B, int kk)
{
  const int total = n * n;
  for(int idx = threadIdx.x; idx < total; idx += blockDim.x)
  {
    const int c = idx / n, r = idx - c * n;
    if(r >= c) D[idx] = D[idx] - dot4_rows(kk, B + r, n, B + c, n);
  }
}

// ---------------------------------------------------------------------------------------------
// Triangular solves on a vector held in shared memory.
// ---------------------------------------------------------------------------------------------

// w(0:n) = L^-1 w, L = n x n lower-triangular view, column-oriented, true division
__device__ __forceinline__ void vec_solve_lower(TView L, int n, double * w)
{
  for(int k = 0; k < n; ++k)
  {
    const double xk = w[k] / L(k, k);
    __syncthreads(); // w[k] has been read by everybody
    for(int i = k + threadIdx.x; i < n; i += blockDim.x) w[i] = i == k ? xk : fma(-xk, L(i, k), w[i]);
    __syncthreads();
  }
}

// w(0:n) = L^-T w
__device__ __forceinline__ void vec_solve_lower_t(TView L, int n, double * w)
{
  for(int k = n - 1; k >= 0; --k)
  {
    const double xk = w[k] / L(k, k);
    __syncthreads();
    for(int i = threadIdx.x; i <= k; i += blockDim.x) w[i] = i == k ? xk : fma(-xk, L(k, i), w[i]);
    __syncthreads();
  }
}

// w(0:rows) -= B x, B = rows x kk view
__device__ __forceinline__ void vec_gemv_sub(double * w, int rows, TView B, int kk, const double * x)
{
  for(int r = threadIdx.x; r < rows; r += blockDim.x) w[r] = w[r] - dot4_rows(kk, B.p + r * B.rs, B.cs, x, 1);
  __syncthreads();
}

// StructuredG::solveL (transpose = false) / solveInPlaceLTranspose (transpose = true) on a vector v that is
// already in shared memory, IN THE PERMUTED NUMBERING for the up arrow (sg_perm: the L solve takes v = P^T m,
// the L^T solve returns m = P v, src/decomposition/blockArrowLLT.cpp:163-169,264-270). start / end are the
// reference's hints in the caller's numbering (end < 0: none). Lt, Bt: two nmax x nmax tiles of scratch.
// Every thread of the CTA must call it, after a barrier that makes v visible; ends on a barrier.
__device__ __forceinline__ int sg_perm(const StructParams & P, int i)
{
  const int n0 = P.size[0];
  return i < n0 ? P.n - n0 + i : i - n0;
}

__device__ __forceinline__ void sg_solve_inplace(const StructParams & P, const double * base, double * v, double * Lt, double * Bt, bool transpose, int start, int end)
{
  const int n = P.n, b = P.b;
  const bool tri = P.type == SG_TRI;
  const bool up = P.type == SG_ARROW_UP;
  const int n0 = P.size[0];
  if(end < 0) end = n;
  if(tri && !transpose)
  {
    // triBlockDiagLSolve (src/decomposition/triBlockDiagLLT.cpp:38-98)
    int nn = 0, l = 0, li = 0;
    bool zero = true;
    for(int i = 0; i < b; ++i)
    {
      const int ni = P.size[i];
      if(nn + ni >= start)
      {
        load_tile(Lt, base + P.doff[i], ni, ni, P.dld[i], false);
        if(zero)
        {
          __syncthreads();
          const int r = nn + ni - start;
          vec_solve_lower(TView{Lt, 1, ni}.sub(ni - r, ni - r), r, v + start);
          zero = false;
        }
        else
        {
          load_tile(Bt, base + P.ooff[i - 1], ni, li, P.old[i - 1], false);
          __syncthreads();
          vec_gemv_sub(v + nn, ni, TView{Bt, 1, ni}, li, v + l);
          vec_solve_lower(TView{Lt, 1, ni}, ni, v + nn);
        }
      }
      l = nn;
      li = ni;
      nn += ni;
    }
  }
  else if(tri)
  {
    // triBlockDiagLTransposeSolve (src/decomposition/triBlockDiagLLT.cpp:100-158)
    int nn = n, l = 0, li = 0;
    bool zero = true;
    for(int i = b - 1; i >= 0; --i)
    {
      const int ni = P.size[i];
      if(nn - ni < end)
      {
        load_tile(Lt, base + P.doff[i], ni, ni, P.dld[i], false);
        if(zero)
        {
          __syncthreads();
          const int r = end - nn + ni;
          vec_solve_lower_t(TView{Lt, 1, ni}, r, v + nn - ni);
          zero = false;
        }
        else
        {
          load_tile(Bt, base + P.ooff[i], li, ni, P.old[i], false); // S_i: n_{i+1} x n_i
          __syncthreads();
          vec_gemv_sub(v + nn - ni, ni, TView{Bt, 1, li}.t(), li, v + l - li);
          vec_solve_lower_t(TView{Lt, 1, ni}, ni, v + nn - ni);
        }
      }
      l = nn;
      li = ni;
      nn -= ni;
    }
  }
  else
  {
    const int last = up ? 0 : b - 1;
    const int nl = P.size[last];
    if(up)
    {
      // hints are given in the caller's row numbering; the permuted system starts n0 rows earlier
      if(!transpose)
      {
        start = max(0, start - n0);
        end = max(0, end - n0);
      }
    }
    double * vb = v + n - nl; // rows of the last block of the permuted system
    if(!transpose)
    {
      // blockArrowLSolve_ (src/decomposition/blockArrowLLT.cpp:92-152)
      int nn = 0;
      for(int i = 0; i < b - 1; ++i)
      {
        const int di = up ? i + 1 : i;
        const int ni = P.size[di];
        const int s = max(start - nn, 0);
        if(ni < s || end <= nn)
        {
          nn += ni;
          continue;
        }
        const int srows = up ? ni : nl, scols = up ? n0 : ni;
        load_tile(Lt, base + P.doff[di], ni, ni, P.dld[di], false);
        load_tile(Bt, base + P.ooff[i], srows, scols, P.old[i], false);
        __syncthreads();
        vec_solve_lower(TView{Lt, 1, ni}.sub(s, s), ni - s, v + nn + s);
        TView B = up ? TView{Bt, 1, srows}.t() : TView{Bt, 1, srows}; // B_i: nl x ni
        vec_gemv_sub(vb, nl, B.sub(0, s), ni - s, v + nn + s);
        nn += ni;
      }
      load_tile(Lt, base + P.doff[last], nl, nl, P.dld[last], false);
      __syncthreads();
      vec_solve_lower(TView{Lt, 1, nl}, nl, vb);
    }
    else
    {
      // blockArrowLTransposeSolve_ (src/decomposition/blockArrowLLT.cpp:176-252)
      bool zero = false;
      if(end > n - nl)
      {
        const int r = end - n + nl;
        load_tile(Lt, base + P.doff[last], nl, nl, P.dld[last], false);
        __syncthreads();
        vec_solve_lower_t(TView{Lt, 1, nl}, r, vb);
      }
      else
        zero = true;
      int nn = 0;
      for(int i = 0; i < b - 1; ++i)
      {
        const int di = up ? i + 1 : i;
        const int ni = P.size[di];
        if(zero && start >= nn + ni)
        {
          nn += ni;
          continue;
        }
        const int srows = up ? ni : nl, scols = up ? n0 : ni;
        __syncthreads();
        if(!zero) load_tile(Bt, base + P.ooff[i], srows, scols, P.old[i], false);
        if(end >= nn) load_tile(Lt, base + P.doff[di], ni, ni, P.dld[di], false);
        __syncthreads();
        if(!zero)
        {
          TView B = up ? TView{Bt, 1, srows}.t() : TView{Bt, 1, srows}; // B_i: nl x ni
          vec_gemv_sub(v + nn, ni, B.t(), nl, vb);
        }
        if(end >= nn) vec_solve_lower_t(TView{Lt, 1, ni}, end >= nn + ni ? ni : end - nn, v + nn);
        nn += ni;
      }
    }
  }
  __syncthreads();
}

} // namespace jrlqp
