// Dense cold-start Goldfarb-Idnani solver: ONE QP PER WARP, persistent work-queue kernel, sm_100a.
//
// Replaces, for a batch of independent QPs, the reference path
//   GoldfarbIdnaniSolver::solve -> DualSolver::solve -> {init_, selectViolatedConstraint_,
//   computeStep_, computeStepLength_, addConstraint_, removeConstraint_}
//   (src/GoldfarbIdnaniSolver.cpp:18-338, src/DualSolver.cpp:38-244, src/internal/ActiveSet.cpp).
//
// Layout (all per-QP state lives in the warp's shared memory for the whole solve; HBM traffic is
// the compulsory input read + output write only):
//   Jb   n x ldj row-major, ldj odd. First holds the lower triangle of G, is factorised in place
//        (L), then J = L^-T is built in its upper triangle and the lower triangle is cleared.
//        Row-major + odd ld makes BOTH access patterns conflict-free: lanes over columns
//        (d = J^T n+) and lanes over rows (z = J2 d2, Givens column rotations).
//   Rp   packed upper-triangular R, column k at k(k+1)/2 (lanes run down a column).
//   xs, zs, ds, rs, us  vectors; alist (int) ordered active list; stat (int8) activation status.
//   Cs   optional staged copy of C (mc x ldcs, ldcs odd, one normal per row).
// Thread mapping: lane l owns rows/columns/constraints l, l+32, ... (RPT = ceil(n/32) slots).
// Every floating-point result is produced in the canonical order that oracle/gi_oracle.cpp
// documents (dot4 / dot32 / fma axpy / Eigen makeGivens), so results are bit-identical to the oracle.
#pragma once

#include "gi_params.h"

#include <cuda_runtime.h>

namespace jrlqp
{

#define JRLQP_FULL 0xffffffffu

template<int RPT>
__device__ __forceinline__ double pick(const double (&v)[RPT], int slot)
{
  double r = v[0];
#pragma unroll
  for(int s = 1; s < RPT; ++s)
    if(slot == s) r = v[s];
  return r;
}

__device__ __forceinline__ double warp_sum32(double acc)
{
  // dot32 butterfly: acc[l] += acc[l ^ off], off = 16,8,4,2,1 (addition commutes => all lanes agree)
#pragma unroll
  for(int off = 16; off >= 1; off >>= 1) acc = acc + __shfl_xor_sync(JRLQP_FULL, acc, off);
  return acc;
}

// Eigen JacobiRotation::makeGivens (real case); same operation order as the oracle.
__device__ __forceinline__ void make_givens(double p, double q, double & c, double & s, double & r)
{
  if(q == 0.0)
  {
    c = p < 0.0 ? -1.0 : 1.0;
    s = 0.0;
    r = fabs(p);
  }
  else if(p == 0.0)
  {
    c = 0.0;
    s = q < 0.0 ? 1.0 : -1.0;
    r = fabs(q);
  }
  else if(fabs(p) > fabs(q))
  {
    double t = q / p;
    double u = sqrt(fma(t, t, 1.0));
    if(p < 0.0) u = -u;
    c = 1.0 / u;
    s = -t * c;
    r = p * u;
  }
  else
  {
    double t = p / q;
    double u = sqrt(fma(t, t, 1.0));
    if(q < 0.0) u = -u;
    s = -1.0 / u;
    c = -t * s;
    r = q * u;
  }
}

// dot4 of two unit-stride vectors, evaluated redundantly by every lane (uniform result).
__device__ __forceinline__ double dot4_uniform(int len, const double * __restrict__ a, const double * __restrict__ b)
{
  double c0 = 0, c1 = 0, c2 = 0, c3 = 0;
  int k = 0;
  for(; k + 3 < len; k += 4)
  {
    c0 = fma(a[k], b[k], c0);
    c1 = fma(a[k + 1], b[k + 1], c1);
    c2 = fma(a[k + 2], b[k + 2], c2);
    c3 = fma(a[k + 3], b[k + 3], c3);
  }
  if(k < len) c0 = fma(a[k], b[k], c0);
  if(k + 1 < len) c1 = fma(a[k + 1], b[k + 1], c1);
  if(k + 2 < len) c2 = fma(a[k + 2], b[k + 2], c2);
  return (c0 + c1) + (c2 + c3);
}

struct Sel
{
  int p;
  int st;
};

// Exact sequential restatement of selectViolatedConstraint_ (src/GoldfarbIdnaniSolver.cpp:84-134),
// every lane redundantly. Only used when some constraint has BOTH slacks negative (bl > bu), the
// one situation where the parallel "first minimum" differs from the reference's else-if chain.
__device__ __noinline__ Sel select_sequential(int n,
                                              int mc,
                                              int nb,
                                              const double * Cbase,
                                              long long ldC,
                                              const double * xs,
                                              const double * bl,
                                              const double * bu,
                                              const double * xl,
                                              const double * xu,
                                              const signed char * stat)
{
  double smin = 0;
  Sel sel{-1, ST_INACTIVE};
  for(int i = 0; i < mc; ++i)
  {
    if(stat[i] == ST_INACTIVE)
    {
      double cx = dot4_uniform(n, Cbase + (long long)i * ldC, xs);
      double sl = cx - bl[i];
      if(sl < smin)
      {
        smin = sl;
        sel = {i, ST_LOWER};
      }
      else
      {
        double su = bu[i] - cx;
        if(su < smin)
        {
          smin = su;
          sel = {i, ST_UPPER};
        }
      }
    }
  }
  for(int i = 0; i < nb; ++i)
  {
    if(stat[mc + i] == ST_INACTIVE)
    {
      double sl = xs[i] - xl[i];
      if(sl < smin)
      {
        smin = sl;
        sel = {mc + i, ST_LOWER_BOUND};
      }
      else
      {
        double su = xu[i] - xs[i];
        if(su < smin)
        {
          smin = su;
          sel = {mc + i, ST_UPPER_BOUND};
        }
      }
    }
  }
  return sel;
}

template<int RPT, bool STAGE_C>
struct GiWarp
{
  // ---- immutable per-launch
  const GiParams & P;
  const int lane;
  const int n, mc, nb, m, ldj;
  double *Jb, *Rp, *xs, *zs, *ds, *rs, *us, *Cs;
  int * alist;
  signed char * stat;
  // ---- per-problem views
  const double *Cb, *bl, *bu, *xl, *xu;
  long long ldC; // leading dimension of the constraint-normal storage Cb (staged or global)
  // ---- solver state (uniform across lanes)
  int q;
  double f;

  __device__ GiWarp(const GiParams & p, double * smem)
  : P(p), lane(threadIdx.x & 31), n(p.n), mc(p.mc), nb(p.nb), m(p.mc + p.nb), ldj(p.ldj)
  {
    Jb = smem;
    Rp = smem + p.off_R;
    xs = smem + p.off_x;
    zs = smem + p.off_z;
    ds = smem + p.off_d;
    rs = smem + p.off_r;
    us = smem + p.off_u;
    Cs = smem + p.off_C;
    alist = reinterpret_cast<int *>(smem + p.off_alist);
    stat = reinterpret_cast<signed char *>(smem + p.off_stat);
  }

  __device__ __forceinline__ int idx(int s) const { return lane + 32 * s; }
  __device__ __forceinline__ static int colR(int k) { return (k * (k + 1)) >> 1; }

  // ------------------------------------------------------------------------------------------
  // init_ (src/GoldfarbIdnaniSolver.cpp:56-82): Cholesky, J = L^-T, x = -G^-1 a, f = a.x/2
  // ------------------------------------------------------------------------------------------
  __device__ bool init(long long b)
  {
    const double * __restrict__ Gb = P.G + b * P.sG;
    const double * __restrict__ ab = P.a + b * P.sa;
    const int ldg = P.ldg;

    // stage the lower triangle of G (column-major in HBM: lanes run down a column => coalesced)
#pragma unroll 4
    for(int j = 0; j < n; ++j)
    {
#pragma unroll
      for(int s = 0; s < RPT; ++s)
      {
        int i = idx(s);
        if(i < n && i >= j) Jb[i * ldj + j] = __ldg(Gb + i + (long long)j * ldg);
      }
    }
    if(STAGE_C)
    {
      // C is n x mc column-major: column i (one normal) is contiguous => coalesced along k
      const double * __restrict__ Cg = P.C + b * P.sC;
      for(int i = 0; i < mc; ++i)
      {
#pragma unroll
        for(int s = 0; s < RPT; ++s)
        {
          int k = idx(s);
          if(k < n) Cs[i * P.ldcs + k] = __ldg(Cg + k + (long long)i * P.ldc);
        }
      }
    }
    __syncwarp();

    // --- left-looking Cholesky, lane = row
    for(int k = 0; k < n; ++k)
    {
      double acc[RPT][4];
#pragma unroll
      for(int s = 0; s < RPT; ++s) acc[s][0] = acc[s][1] = acc[s][2] = acc[s][3] = 0.0;
      const double * Lk = Jb + k * ldj;
      const double * Li[RPT];
#pragma unroll
      for(int s = 0; s < RPT; ++s) Li[s] = Jb + min(idx(s), n - 1) * ldj;
      int j = 0;
      for(; j + 3 < k; j += 4)
      {
        double l0 = Lk[j], l1 = Lk[j + 1], l2 = Lk[j + 2], l3 = Lk[j + 3];
#pragma unroll
        for(int s = 0; s < RPT; ++s)
        {
          if(32 * s + 31 < k) continue; // whole slot above the pivot row: nothing to do
          acc[s][0] = fma(Li[s][j], l0, acc[s][0]);
          acc[s][1] = fma(Li[s][j + 1], l1, acc[s][1]);
          acc[s][2] = fma(Li[s][j + 2], l2, acc[s][2]);
          acc[s][3] = fma(Li[s][j + 3], l3, acc[s][3]);
        }
      }
#pragma unroll
      for(int t = 0; t < 3; ++t)
      {
        if(j + t < k)
        {
          double lt = Lk[j + t];
#pragma unroll
          for(int s = 0; s < RPT; ++s) acc[s][t] = fma(Li[s][j + t], lt, acc[s][t]);
        }
      }
      double v[RPT];
#pragma unroll
      for(int s = 0; s < RPT; ++s) v[s] = Li[s][k] - ((acc[s][0] + acc[s][1]) + (acc[s][2] + acc[s][3]));
      double vk = __shfl_sync(JRLQP_FULL, pick<RPT>(v, k >> 5), k & 31);
      if(vk <= 0.0) return false; // Eigen llt: "if (x <= 0) return k" -> NON_POS_HESSIAN
      double lkk = sqrt(vk);
      __syncwarp();
#pragma unroll
      for(int s = 0; s < RPT; ++s)
      {
        int i = idx(s);
        if(i == k)
          Jb[i * ldj + k] = lkk;
        else if(i > k && i < n)
          Jb[i * ldj + k] = v[s] / lkk;
      }
      __syncwarp();
    }

    // --- x = -G^-1 a : column-oriented forward and backward substitution, lane = row
    double y[RPT];
#pragma unroll
    for(int s = 0; s < RPT; ++s) y[s] = idx(s) < n ? __ldg(ab + idx(s)) : 0.0;
    for(int k = 0; k < n; ++k)
    {
      double yk = __shfl_sync(JRLQP_FULL, pick<RPT>(y, k >> 5), k & 31) / Jb[k * ldj + k];
#pragma unroll
      for(int s = 0; s < RPT; ++s)
      {
        int i = idx(s);
        if(i == k)
          y[s] = yk;
        else if(i > k && i < n)
          y[s] = fma(-yk, Jb[i * ldj + k], y[s]);
      }
    }
    for(int k = n - 1; k >= 0; --k)
    {
      double xk = __shfl_sync(JRLQP_FULL, pick<RPT>(y, k >> 5), k & 31) / Jb[k * ldj + k];
#pragma unroll
      for(int s = 0; s < RPT; ++s)
      {
        int i = idx(s);
        if(i == k)
          y[s] = xk;
        else if(i < k)
          y[s] = fma(-xk, Jb[k * ldj + i], y[s]);
      }
    }
    double facc = 0.0;
#pragma unroll
    for(int s = 0; s < RPT; ++s)
    {
      int i = idx(s);
      if(i < n)
      {
        double xi = -y[s];
        xs[i] = xi;
        facc = fma(__ldg(ab + i), xi, facc);
      }
    }
    f = 0.5 * warp_sum32(facc);

    // --- reciprocals of the diagonal, optional copy-out of L (what the reference leaves in G)
#pragma unroll
    for(int s = 0; s < RPT; ++s)
    {
      int i = idx(s);
      if(i < n) rs[i] = 1.0 / Jb[i * ldj + i];
    }
    if(P.L != nullptr)
    {
      double * Lout = P.L + b * (long long)n * n;
      for(int j = 0; j < n; ++j)
      {
#pragma unroll
        for(int s = 0; s < RPT; ++s)
        {
          int i = idx(s);
          if(i < n && i >= j) Lout[i + (long long)j * n] = Jb[i * ldj + j];
        }
      }
    }
    __syncwarp();

    // --- J = L^-T in place (upper triangle), lane = column
#pragma unroll
    for(int s = 0; s < RPT; ++s)
    {
      int j = idx(s);
      if(j < n) Jb[j * ldj + j] = rs[j];
    }
    __syncwarp();
    for(int i = n - 2; i >= 0; --i)
    {
      double acc[RPT][4];
#pragma unroll
      for(int s = 0; s < RPT; ++s) acc[s][0] = acc[s][1] = acc[s][2] = acc[s][3] = 0.0;
      for(int k0 = i + 1; k0 < n; k0 += 4)
      {
#pragma unroll
        for(int t = 0; t < 4; ++t)
        {
          int k = k0 + t;
          if(k < n)
          {
            double lki = Jb[k * ldj + i];
#pragma unroll
            for(int s = 0; s < RPT; ++s)
            {
              int j = idx(s);
              if(k <= j && j < n) acc[s][t] = fma(lki, Jb[k * ldj + j], acc[s][t]);
            }
          }
        }
      }
      double ri = rs[i];
      __syncwarp();
#pragma unroll
      for(int s = 0; s < RPT; ++s)
      {
        int j = idx(s);
        if(j > i && j < n) Jb[i * ldj + j] = (-((acc[s][0] + acc[s][1]) + (acc[s][2] + acc[s][3]))) * ri;
      }
      __syncwarp();
    }
    // clear the strict lower triangle (L is no longer needed)
    for(int r = 1; r < n; ++r)
    {
#pragma unroll
      for(int s = 0; s < RPT; ++s)
      {
        int c = idx(s);
        if(c < r) Jb[r * ldj + c] = 0.0;
      }
    }
    // A_.reset()
    for(int i = lane; i < m; i += 32) stat[i] = ST_INACTIVE;
    q = 0;
    __syncwarp();
    return true;
  }

  // ------------------------------------------------------------------------------------------
  // selectViolatedConstraint_ (src/GoldfarbIdnaniSolver.cpp:84-134), lane = constraint.
  // Also returns the selected constraint's cx so that computeStepLength_ can reuse it (x is
  // unchanged between the two when step 1 was executed; same dot4 => same bits).
  // ------------------------------------------------------------------------------------------
  __device__ Sel select(double & cx_sel)
  {
    double best = 0.0;
    double bestcx = 0.0;
    int code = 0x7fffffff;
    bool bothneg = false;
    for(int base = 0; base < mc; base += 32)
    {
      int i = base + lane;
      bool act = i < mc && stat[i] == ST_INACTIVE;
      if(__ballot_sync(JRLQP_FULL, act) == 0u) continue;
      const double * ci = Cb + (long long)min(i, mc - 1) * ldC;
      double c0 = 0, c1 = 0, c2 = 0, c3 = 0;
      int k = 0;
      for(; k + 3 < n; k += 4)
      {
        c0 = fma(ci[k], xs[k], c0);
        c1 = fma(ci[k + 1], xs[k + 1], c1);
        c2 = fma(ci[k + 2], xs[k + 2], c2);
        c3 = fma(ci[k + 3], xs[k + 3], c3);
      }
      if(k < n) c0 = fma(ci[k], xs[k], c0);
      if(k + 1 < n) c1 = fma(ci[k + 1], xs[k + 1], c1);
      if(k + 2 < n) c2 = fma(ci[k + 2], xs[k + 2], c2);
      double cx = (c0 + c1) + (c2 + c3);
      if(act)
      {
        double sl = cx - bl[i];
        double su = bu[i] - cx;
        if(sl < 0.0 && su < 0.0) bothneg = true;
        if(sl < best)
        {
          best = sl;
          bestcx = cx;
          code = i * 8 + ST_LOWER;
        }
        else if(su < best)
        {
          best = su;
          bestcx = cx;
          code = i * 8 + ST_UPPER;
        }
      }
    }
    for(int base = 0; base < nb; base += 32)
    {
      int i = base + lane;
      if(i < nb && stat[mc + i] == ST_INACTIVE)
      {
        double xi = xs[i];
        double sl = xi - xl[i];
        double su = xu[i] - xi;
        if(sl < 0.0 && su < 0.0) bothneg = true;
        if(sl < best)
        {
          best = sl;
          bestcx = xi;
          code = (mc + i) * 8 + ST_LOWER_BOUND;
        }
        else if(su < best)
        {
          best = su;
          bestcx = xi;
          code = (mc + i) * 8 + ST_UPPER_BOUND;
        }
      }
    }
    if(__any_sync(JRLQP_FULL, bothneg))
    {
      Sel s = select_sequential(n, mc, nb, Cb, ldC, xs, bl, bu, xl, xu, stat);
      cx_sel = s.p < 0 ? 0.0 : (s.p < mc ? dot4_uniform(n, Cb + (long long)s.p * ldC, xs) : xs[s.p - mc]);
      return s;
    }
#pragma unroll
    for(int off = 16; off >= 1; off >>= 1)
    {
      double ov = __shfl_xor_sync(JRLQP_FULL, best, off);
      double ocx = __shfl_xor_sync(JRLQP_FULL, bestcx, off);
      int oc = __shfl_xor_sync(JRLQP_FULL, code, off);
      if(ov < best || (ov == best && oc < code))
      {
        best = ov;
        bestcx = ocx;
        code = oc;
      }
    }
    code = __shfl_sync(JRLQP_FULL, code, 0); // keep control flow uniform even with NaN inputs
    cx_sel = __shfl_sync(JRLQP_FULL, bestcx, 0);
    if(code == 0x7fffffff) return {-1, ST_INACTIVE};
    return {code >> 3, code & 7};
  }

  // ------------------------------------------------------------------------------------------
  // computeStep_ (src/GoldfarbIdnaniSolver.cpp:136-148): d = J^T n+, z = J2 d2, r = R^-1 d1.
  // Leaves d in ds, z in zs (and in zreg), r in rs.
  // ------------------------------------------------------------------------------------------
  __device__ void compute_step(Sel sc, double (&zreg)[RPT])
  {
    // d, lane = column
    if(sc.st < ST_LOWER_BOUND)
    {
      const double * __restrict__ c = Cb + (long long)sc.p * ldC;
      double acc[RPT][4];
#pragma unroll
      for(int s = 0; s < RPT; ++s) acc[s][0] = acc[s][1] = acc[s][2] = acc[s][3] = 0.0;
      int col[RPT];
#pragma unroll
      for(int s = 0; s < RPT; ++s) col[s] = min(idx(s), n - 1);
      int i = 0;
      for(; i + 3 < n; i += 4)
      {
        double c0 = c[i], c1 = c[i + 1], c2 = c[i + 2], c3 = c[i + 3];
        const double * Ji = Jb + i * ldj;
#pragma unroll
        for(int s = 0; s < RPT; ++s)
        {
          acc[s][0] = fma(Ji[col[s]], c0, acc[s][0]);
          acc[s][1] = fma(Ji[ldj + col[s]], c1, acc[s][1]);
          acc[s][2] = fma(Ji[2 * ldj + col[s]], c2, acc[s][2]);
          acc[s][3] = fma(Ji[3 * ldj + col[s]], c3, acc[s][3]);
        }
      }
#pragma unroll
      for(int t = 0; t < 3; ++t)
      {
        if(i + t < n)
        {
          double ct = c[i + t];
#pragma unroll
          for(int s = 0; s < RPT; ++s) acc[s][t] = fma(Jb[(i + t) * ldj + col[s]], ct, acc[s][t]);
        }
      }
#pragma unroll
      for(int s = 0; s < RPT; ++s)
      {
        int j = idx(s);
        double dj = (acc[s][0] + acc[s][1]) + (acc[s][2] + acc[s][3]);
        if(sc.st == ST_UPPER) dj = -dj;
        if(j < n) ds[j] = dj;
      }
    }
    else
    {
      const double * Jrow = Jb + (sc.p - mc) * ldj;
#pragma unroll
      for(int s = 0; s < RPT; ++s)
      {
        int j = idx(s);
        if(j < n) ds[j] = sc.st == ST_UPPER_BOUND ? -Jrow[j] : Jrow[j];
      }
    }
    __syncwarp();

    // z, lane = row: z[i] = dot4_{j=q..n-1}(J(i,j), d[j]), accumulator (j-q)&3
    {
      double acc[RPT][4];
#pragma unroll
      for(int s = 0; s < RPT; ++s) acc[s][0] = acc[s][1] = acc[s][2] = acc[s][3] = 0.0;
      const double * Jr[RPT];
#pragma unroll
      for(int s = 0; s < RPT; ++s) Jr[s] = Jb + min(idx(s), n - 1) * ldj;
      int j = q;
      for(; j + 3 < n; j += 4)
      {
        double d0 = ds[j], d1 = ds[j + 1], d2 = ds[j + 2], d3 = ds[j + 3];
#pragma unroll
        for(int s = 0; s < RPT; ++s)
        {
          acc[s][0] = fma(Jr[s][j], d0, acc[s][0]);
          acc[s][1] = fma(Jr[s][j + 1], d1, acc[s][1]);
          acc[s][2] = fma(Jr[s][j + 2], d2, acc[s][2]);
          acc[s][3] = fma(Jr[s][j + 3], d3, acc[s][3]);
        }
      }
#pragma unroll
      for(int t = 0; t < 3; ++t)
      {
        if(j + t < n)
        {
          double dt = ds[j + t];
#pragma unroll
          for(int s = 0; s < RPT; ++s) acc[s][t] = fma(Jr[s][j + t], dt, acc[s][t]);
        }
      }
#pragma unroll
      for(int s = 0; s < RPT; ++s)
      {
        int i = idx(s);
        zreg[s] = (acc[s][0] + acc[s][1]) + (acc[s][2] + acc[s][3]);
        if(i < n)
          zs[i] = zreg[s];
        else
          zreg[s] = 0.0;
      }
    }

    // r = R^-1 d(0:q): column-oriented back substitution, lane = row, true division
    {
      double w[RPT], rr[RPT];
#pragma unroll
      for(int s = 0; s < RPT; ++s)
      {
        w[s] = idx(s) < q ? ds[idx(s)] : 0.0;
        rr[s] = 0.0;
      }
      for(int k = q - 1; k >= 0; --k)
      {
        const double * Rk = Rp + colR(k);
        double rk = __shfl_sync(JRLQP_FULL, pick<RPT>(w, k >> 5), k & 31) / Rk[k];
#pragma unroll
        for(int s = 0; s < RPT; ++s)
        {
          int j = idx(s);
          if(j == k)
            rr[s] = rk;
          else if(j < k)
            w[s] = fma(-rk, Rk[j], w[s]);
        }
      }
#pragma unroll
      for(int s = 0; s < RPT; ++s)
        if(idx(s) < q) rs[idx(s)] = rr[s];
    }
    __syncwarp();
  }

  // ------------------------------------------------------------------------------------------
  // computeStepLength_ (src/GoldfarbIdnaniSolver.cpp:150-219), incl. the activationStatus(k) quirk.
  // nz receives ConstraintNormal::dot(z) (src/GoldfarbIdnaniSolver.cpp:289-293).
  // ------------------------------------------------------------------------------------------
  __device__ void step_length(Sel sc,
                              const double (&zreg)[RPT],
                              bool cx_valid,
                              double cx_in,
                              bool partial_only_t2,
                              double & t1,
                              double & t2,
                              int & l,
                              double & nz)
  {
    (void)partial_only_t2;
    const double big = P.big_bnd;
    // t1: first minimum of u[k]/r[k] over r[k] > 0 and status_[k] not in {EQUALITY, FIXED}
    double bt = big;
    int bl_ = 0x7fffffff;
#pragma unroll
    for(int s = 0; s < RPT; ++s)
    {
      int k = idx(s);
      if(k < q)
      {
        int sk = stat[k]; // NOTE: indexed by the position k, as in the reference (quirk)
        double rk = rs[k];
        if(sk != ST_EQUALITY && sk != ST_FIXED && rk > 0.0)
        {
          double tk = us[k] / rk;
          if(tk < bt)
          {
            bt = tk;
            bl_ = k;
          }
        }
      }
    }
#pragma unroll
    for(int off = 16; off >= 1; off >>= 1)
    {
      double ot = __shfl_xor_sync(JRLQP_FULL, bt, off);
      int ol = __shfl_xor_sync(JRLQP_FULL, bl_, off);
      if(ot < bt || (ot == bt && ol < bl_))
      {
        bt = ot;
        bl_ = ol;
      }
    }
    bl_ = __shfl_sync(JRLQP_FULL, bl_, 0);
    bt = __shfl_sync(JRLQP_FULL, bt, 0);
    t1 = bl_ == 0x7fffffff ? big : bt;
    l = bl_ == 0x7fffffff ? 0 : bl_;

    // ||z||
    double zz = 0.0;
#pragma unroll
    for(int s = 0; s < RPT; ++s) zz = fma(zreg[s], zreg[s], zz);
    double znorm = sqrt(warp_sum32(zz));

    t2 = big;
    double cz;
    if(sc.st < ST_LOWER_BOUND)
    {
      const double * __restrict__ c = Cb + (long long)sc.p * ldC;
      cz = dot4_uniform(n, c, zs);
      nz = sc.st == ST_UPPER ? -cz : cz;
      if(znorm > 1e-14)
      {
        double b = sc.st == ST_UPPER ? bu[sc.p] : bl[sc.p]; // EQUALITY: bl (addInitialConstraint)
        double cx = cx_valid ? cx_in : dot4_uniform(n, c, xs);
        t2 = (b - cx) / cz;
      }
    }
    else
    {
      int pb = sc.p - mc;
      cz = zs[pb];
      nz = sc.st == ST_UPPER_BOUND ? -cz : cz;
      if(znorm > 1e-14)
      {
        double b = sc.st == ST_UPPER_BOUND ? xu[pb] : xl[pb];
        t2 = (b - xs[pb]) / cz;
      }
    }
  }

  // x += t z ; f += t (n+.z) (t/2 + u[q]) ; u(0:q) -= t r ; u[q] += t
  __device__ void take_step(double t, double nz, const double (&zreg)[RPT], bool primal)
  {
    double uq = us[q];
    if(primal)
    {
#pragma unroll
      for(int s = 0; s < RPT; ++s)
      {
        int i = idx(s);
        if(i < n) xs[i] = fma(t, zreg[s], xs[i]);
      }
      f += (t * nz) * (0.5 * t + uq);
    }
    __syncwarp();
#pragma unroll
    for(int s = 0; s < RPT; ++s)
    {
      int k = idx(s);
      if(k < q) us[k] = fma(-t, rs[k], us[k]);
    }
    if(lane == 0) us[q] = uq + t;
    __syncwarp();
  }

  // ------------------------------------------------------------------------------------------
  // addConstraint (src/DualSolver.cpp:231-235) + addConstraint_ (src/GoldfarbIdnaniSolver.cpp:221-237)
  // ------------------------------------------------------------------------------------------
  __device__ void add_constraint(Sel sc)
  {
    if(lane == 0)
    {
      alist[q] = sc.p;
      stat[sc.p] = (signed char)sc.st;
    }
    q += 1;
    // Givens sweep i = n-2 .. q-1 on (d[i], d[i+1]); lane = row of J, the rotated value of
    // column i+1 is carried in a register from one rotation to the next.
    double yprev[RPT];
    double * Jr[RPT];
#pragma unroll
    for(int s = 0; s < RPT; ++s)
    {
      Jr[s] = Jb + min(idx(s), n - 1) * ldj;
      yprev[s] = Jr[s][n - 1];
    }
    double rho = ds[n - 1];
    for(int i = n - 2; i >= q - 1; --i)
    {
      double c, sn, r;
      make_givens(ds[i], rho, c, sn, r);
      rho = r;
#pragma unroll
      for(int s = 0; s < RPT; ++s)
      {
        double xi = Jr[s][i];
        double yi = yprev[s];
        if(idx(s) < n) Jr[s][i + 1] = fma(c, yi, sn * xi);
        yprev[s] = fma(c, xi, -(sn * yi));
      }
    }
    if(q - 1 <= n - 2)
    {
#pragma unroll
      for(int s = 0; s < RPT; ++s)
        if(idx(s) < n) Jr[s][q - 1] = yprev[s];
    }
    // R(0:q, q-1) = d(0:q), with d[q-1] = rho
    double * Rq = Rp + colR(q - 1);
#pragma unroll
    for(int s = 0; s < RPT; ++s)
    {
      int k = idx(s);
      if(k < q - 1)
        Rq[k] = ds[k];
      else if(k == q - 1)
        Rq[k] = rho;
    }
    __syncwarp();
  }

  // ------------------------------------------------------------------------------------------
  // removeConstraint (src/DualSolver.cpp:237-244) + removeConstraint_ (src/GoldfarbIdnaniSolver.cpp:239-256)
  // ------------------------------------------------------------------------------------------
  __device__ void remove_constraint(int l)
  {
    // u.segment(l, q-l) = u.tail(q-l) (u has q+1 entries) ; A_.deactivate(l)
    double ut[RPT];
    int at[RPT];
#pragma unroll
    for(int s = 0; s < RPT; ++s)
    {
      int k = idx(s);
      ut[s] = (k >= l && k < q) ? us[k + 1] : 0.0;
      at[s] = (k >= l && k + 1 < q) ? alist[k + 1] : -1;
    }
    int removed = alist[l];
    __syncwarp();
#pragma unroll
    for(int s = 0; s < RPT; ++s)
    {
      int k = idx(s);
      if(k >= l && k < q) us[k] = ut[s];
      if(k >= l && k + 1 < q) alist[k] = at[s];
    }
    if(lane == 0) stat[removed] = ST_INACTIVE;
    q -= 1;
    __syncwarp();

    for(int i = l; i < q; ++i)
    {
      double * Ri = Rp + colR(i);
      double * Ri1 = Rp + colR(i + 1);
      // R.col(i).head(i) = R.col(i+1).head(i)
#pragma unroll
      for(int s = 0; s < RPT; ++s)
      {
        int k = idx(s);
        if(k < i) Ri[k] = Ri1[k];
      }
      double c, sn, r;
      make_givens(Ri1[i], Ri1[i + 1], c, sn, r);
      __syncwarp();
      if(lane == 0) Ri[i] = r;
      // rows i, i+1 of columns i+2 .. q (lane = column)
#pragma unroll
      for(int s = 0; s < RPT; ++s)
      {
        int j = i + 2 + idx(s);
        if(j <= q)
        {
          double * Rj = Rp + colR(j);
          double xi = Rj[i], yi = Rj[i + 1];
          Rj[i] = fma(c, xi, -(sn * yi));
          Rj[i + 1] = fma(c, yi, sn * xi);
        }
      }
      // columns i, i+1 of J (lane = row)
#pragma unroll
      for(int s = 0; s < RPT; ++s)
      {
        int rrow = idx(s);
        if(rrow < n)
        {
          double * Jr = Jb + rrow * ldj;
          double xi = Jr[i], yi = Jr[i + 1];
          Jr[i] = fma(c, xi, -(sn * yi));
          Jr[i + 1] = fma(c, yi, sn * xi);
        }
      }
      __syncwarp();
    }
  }

  // ------------------------------------------------------------------------------------------
  // initActiveSet / addInitialConstraint (src/GoldfarbIdnaniSolver.cpp:268-338)
  // ------------------------------------------------------------------------------------------
  __device__ void pre_activate(Sel sc)
  {
    if(lane == 0) us[q] = 0.0;
    __syncwarp();
    double zreg[RPT];
    compute_step(sc, zreg);
    double t1, t2, nz;
    int l;
    step_length(sc, zreg, false, 0.0, true, t1, t2, l, nz);
    // t = 0 unless ||z|| > 1e-14 (then the exact step onto the constraint): step_length leaves
    // t2 = bigBnd in the first case.
    double t = t2 == P.big_bnd ? 0.0 : t2;
    take_step(t, nz, zreg, true);
    add_constraint(sc);
  }

  __device__ void init_active_set()
  {
    for(int base = 0; base < mc; base += 32)
    {
      int i = base + lane;
      unsigned eq = __ballot_sync(JRLQP_FULL, i < mc && bl[i] == bu[i]);
      while(eq)
      {
        int bit = __ffs(eq) - 1;
        eq &= eq - 1;
        pre_activate({base + bit, ST_EQUALITY});
      }
    }
    for(int base = 0; base < nb; base += 32)
    {
      int i = base + lane;
      unsigned eq = __ballot_sync(JRLQP_FULL, i < nb && xl[i] == xu[i]);
      while(eq)
      {
        int bit = __ffs(eq) - 1;
        eq &= eq - 1;
        pre_activate({mc + base + bit, ST_FIXED});
      }
    }
  }

  // ------------------------------------------------------------------------------------------
  // DualSolver::solve (src/DualSolver.cpp:91-168) for problem b; writes all outputs.
  // ------------------------------------------------------------------------------------------
  __device__ void solve(long long b)
  {
    bl = P.bl + b * P.sbl;
    bu = P.bu + b * P.sbu;
    xl = nb ? P.xl + b * P.sxl : nullptr;
    xu = nb ? P.xu + b * P.sxu : nullptr;
    if(STAGE_C)
    {
      Cb = Cs;
      ldC = P.ldcs;
    }
    else
    {
      Cb = P.C + b * P.sC;
      ldC = P.ldc;
    }

    int status = TS_MAX_ITER_REACHED;
    int it = 0;
    if(!init(b))
    {
      status = TS_NON_POS_HESSIAN;
      write_failure(b, status);
      return;
    }
    init_active_set();

    bool skip = false;
    Sel sc{-1, ST_INACTIVE};
    double cx_sel = 0.0;
    const double big = P.big_bnd;
    for(; it < P.max_iter; ++it)
    {
      if(!skip)
      {
        sc = select(cx_sel);
        if(sc.st == ST_INACTIVE)
        {
          status = TS_SUCCESS;
          break;
        }
        if(lane == 0) us[q] = 0.0;
        __syncwarp();
      }
      double zreg[RPT];
      compute_step(sc, zreg);
      double t1, t2, nz;
      int l;
      step_length(sc, zreg, !skip, cx_sel, false, t1, t2, l, nz);
      double t = t2 < t1 ? t2 : t1; // std::min(t1, t2)
      if(t >= big)
      {
        status = TS_INFEASIBLE;
        break;
      }
      if(t2 >= big)
      {
        take_step(t, nz, zreg, false);
        remove_constraint(l);
        skip = true;
      }
      else
      {
        take_step(t, nz, zreg, true);
        if(t == t2)
        {
          add_constraint(sc);
          skip = false;
        }
        else
        {
          remove_constraint(l);
          skip = true;
        }
      }
    }
    write_result(b, status, it);
  }

  __device__ void write_result(long long b, int status, int it)
  {
    double * xo = P.x + b * n;
    for(int i = lane; i < n; i += 32) xo[i] = xs[i];
    if(P.u)
    {
      // DualSolver::multipliers (src/DualSolver.cpp:38-69)
      double * uo = P.u + b * m;
      for(int i = lane; i < m; i += 32) uo[i] = 0.0;
      __syncwarp();
      for(int k = lane; k < q; k += 32)
      {
        int i = alist[k];
        int s = stat[i];
        uo[i] = (s == ST_UPPER || s == ST_UPPER_BOUND) ? us[k] : -us[k];
      }
    }
    if(P.active_set)
    {
      signed char * ao = P.active_set + b * m;
      for(int i = lane; i < m; i += 32) ao[i] = stat[i];
    }
    if(P.active_list)
    {
      int * lo = P.active_list + b * n;
      for(int k = lane; k < n; k += 32) lo[k] = k < q ? alist[k] : -1;
    }
    if(lane == 0)
    {
      if(P.f) P.f[b] = f;
      if(P.iterations) P.iterations[b] = it;
      if(P.status) P.status[b] = status;
      if(P.n_active) P.n_active[b] = q;
    }
    __syncwarp();
  }

  __device__ void write_failure(long long b, int status)
  {
    double * xo = P.x + b * n;
    for(int i = lane; i < n; i += 32) xo[i] = 0.0;
    if(P.u)
      for(int i = lane; i < m; i += 32) P.u[b * m + i] = 0.0;
    if(P.active_set)
      for(int i = lane; i < m; i += 32) P.active_set[b * m + i] = ST_INACTIVE;
    if(P.active_list)
      for(int k = lane; k < n; k += 32) P.active_list[b * n + k] = -1;
    if(lane == 0)
    {
      if(P.f) P.f[b] = 0.0;
      if(P.iterations) P.iterations[b] = 0;
      if(P.status) P.status[b] = status;
      if(P.n_active) P.n_active[b] = 0;
    }
    __syncwarp();
  }
};

// Persistent kernel: grid = resident warps of the whole GPU; every warp pulls the next problem
// index from an atomic ticket counter, which absorbs the divergent iteration counts across QPs.
template<int RPT, bool STAGE_C>
__global__ void __launch_bounds__(32) gi_dense_warp_kernel(const GiParams p)
{
  extern __shared__ __align__(16) double smem[];
  GiWarp<RPT, STAGE_C> w(p, smem);
  for(;;)
  {
    unsigned long long b = 0;
    if(w.lane == 0) b = atomicAdd(p.counter, 1ull);
    b = __shfl_sync(JRLQP_FULL, b, 0);
    if(b >= (unsigned long long)p.batch) break;
    w.solve((long long)b);
    __syncwarp();
  }
}

} // namespace jrlqp
