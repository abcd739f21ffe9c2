// Dense cold-start Goldfarb-Idnani solver: ONE QP PER CTA of W warps (T = 32 W >= n threads, one
// row / column per thread), persistent work-queue kernel, sm_100a.
//
// Replaces, for a batch of independent QPs, the reference path
//   GoldfarbIdnaniSolver::solve -> DualSolver::solve -> {init_, selectViolatedConstraint_,
//   computeStep_, computeStepLength_, addConstraint_, removeConstraint_}
//   (src/GoldfarbIdnaniSolver.cpp:18-338, src/DualSolver.cpp:38-244, src/internal/ActiveSet.cpp).
//
// Why this shape (profiles/r01a_*): the algorithm is a chain of short BLAS-2 steps glued by two
// strictly serial scalar recurrences per iteration — the back substitution r = R^-1 d1 (one FP64
// division per link) and the Givens sweep of addConstraint_ (division + square root per link). A
// single warp per QP spends ~8 cycles per instruction on those chains. Here
//   * every BLAS-2 phase is spread over all T threads (thread = row, column or constraint),
//   * the two recurrences, which are independent of each other, run CONCURRENTLY on different
//     warps: warp 0 does the back substitution while warp 1 runs the Givens recurrence
//     speculatively (its rotations are only applied if the step turns out to be a full step),
//   * the Givens recurrence keeps only what is truly serial (t, u, rho); the second division and
//     the products giving (c, s) are done lane-parallel afterwards, and the rotations are applied
//     to J by all threads from the (c, s) table.
//
// Layout (per-QP state lives in the CTA's shared memory for the whole solve; HBM traffic is the
// compulsory input read + output write only):
//   Jb   n x ldj row-major, ldj odd: lower triangle of G -> L (in place) -> J = L^-T in the upper
//        triangle, lower cleared. Row-major + odd ld makes both access patterns conflict-free:
//        threads over columns (d = J^T n+) and threads over rows (z = J2 d2, column rotations).
//   Rp   packed upper-triangular R, column k at k(k+1)/2.
//   xs, zs, ds, rs, us, cv (selected normal), grec (Givens sweep records), ldiag, alist, stat, scratch.
//   Cs   optional staged copy of C (mc x ldcs, ldcs odd, one normal per row).
// Every floating-point result is produced in the canonical order that oracle/gi_oracle.cpp documents
// (dot4 / dot32 / fma axpy / Eigen makeGivens): results are bit-identical to the oracle whatever W.
#pragma once

#include "fp64_exact.cuh"
#include "gi_params.h"

#include <cuda_runtime.h>

namespace jrlqp
{

#define JRLQP_FULL 0xffffffffu
#ifndef JRLQP_DMMA_CHOL
#  define JRLQP_DMMA_CHOL 1 // four-warp kernels: the inner products of the Cholesky over the columns left of the current 16-column
                           // block on the FP64 tensor cores (mma.sync.m8n8k4.f64), bit-identical to the dot4 chains
#endif
#ifndef JRLQP_UNR_CHOL
#  define JRLQP_UNR_CHOL 4
#endif
#ifndef JRLQP_UNR_JB
#  define JRLQP_UNR_JB 4
#endif
#ifndef JRLQP_CT_ALLW
#  define JRLQP_CT_ALLW 0 // 1: compile the transposed-copy scan into the narrow kernels too (tuning comparison)
#endif
#ifndef JRLQP_OPT_PN
#  define JRLQP_OPT_PN 1
#endif
#ifndef JRLQP_OPT_BSPLIT
#  define JRLQP_OPT_BSPLIT 1
#endif
#ifndef JRLQP_OPT_V2_SCAN
#  define JRLQP_OPT_V2_SCAN 0
#endif
#ifndef JRLQP_BS_PREFETCH
#  define JRLQP_BS_PREFETCH 3 // back substitution: operands loaded one link ahead (two links per trip) from this many warps per QP on
#endif
#ifndef JRLQP_CHAIN_VOTE
#  define JRLQP_CHAIN_VOTE 1 // Givens recurrence: the branch to the slow link is taken on a warp vote (no convergence barrier on the fast path)
#endif
#ifndef JRLQP_CHAIN_UNR
#  define JRLQP_CHAIN_UNR 1 // links per trip of the Givens recurrence in the narrow kernels (the wide ones always run 2); 2: -3 % at n = 50 (profiles/r02j_ab_A.txt)
#endif
#ifndef JRLQP_BULK_PREFETCH
#  define JRLQP_BULK_PREFETCH 0 // 1: ticket look-ahead, the next problem's G and C pulled into L2 by TMA bulk prefetches (cp.async.bulk.prefetch.L2)
                                // while the current one is solved. Measured (profiles/r02u_ab_*.txt): -0.4 % at n = 50 and n = 128, -2.8 % at n = 20 — the
                                // staging loads are not what these kernels wait for (long_scoreboard 4 % of the stall samples): off
#endif
#ifndef JRLQP_ZPART_W1
#  define JRLQP_ZPART_W1 1 // (+1.3 % at n = 50, profiles/r03a_ab_A.txt) W == 2: in the main iterations (short Givens chain, long back substitution) the z-dependent half of the step length runs on warp 1
#endif
#ifndef JRLQP_CHAIN_BF
#  define JRLQP_CHAIN_BF 1 // W = 4: the links of the Givens recurrence below the last one the seeds predict on another branch of
                          // makeGivens run in a branch-free loop (operands two links ahead, exact branch test accumulated and checked once)
#endif
#ifndef JRLQP_CHAIN_BF_MINW
#  define JRLQP_CHAIN_BF_MINW 4 // (W = 3, register-capped: -1.2 %, profiles/r4i_ab_bf.txt; W = 2: profiles/r4k_ab_bf_w2.txt)
#endif
#ifndef JRLQP_CHAIN_BF_UNR
#  define JRLQP_CHAIN_BF_UNR 2
#endif
#ifndef JRLQP_SEEDS_IN_Z
#  define JRLQP_SEEDS_IN_Z 1 // (+1.6 % at n = 50, neutral at n = 128, profiles/r03b_ab_*.txt) W > 1: the seeds of the Givens recurrence are computed by ALL the threads next to z = J2 d2 (one element each)
                             // instead of by the chain warp in chunks of 32 before its recurrence
#endif
#ifndef JRLQP_MINB3
#  define JRLQP_MINB3 1 // resident CTAs per SM the three-warp kernel is compiled for (4: 168 registers, 12 warps per SM)
#endif
#ifndef JRLQP_WARM_MINB1
#  define JRLQP_WARM_MINB1 16 // resident CTAs per SM the one-warp WARM kernel is compiled for: 128 registers, no spill, 16 CTAs per SM
                              // instead of 225 registers / 8 CTAs: sequences of n = 20 11.86 M -> 13.33 M QP-steps/s (profiles/r5e_*)
#endif
#ifndef JRLQP_WARM_MINB2
#  define JRLQP_WARM_MINB2 6 // the same for the two-warp WARM kernel (n = 33 ... 64): 160 registers instead of 221, no spill; where shared
                             // memory leaves room (n <= 45) six QPs are resident instead of four: sequences of n = 40 4.30 M -> 5.25 M
                             // QP-steps/s, n = 50 (four QPs per SM by shared memory either way) 3.05 M -> 3.07 M (profiles/r5f_*)
#endif
#ifndef JRLQP_MINB1
#  define JRLQP_MINB1 16 // resident CTAs per SM the one-warp kernel is compiled for (register cap 65536 / (32 * MINB1))
#endif

// Optional per-phase cycle accounting (build with -DJRLQP_PHASE_TIMING; scripts/phase_timing.py):
// every warp adds the cycles it spends in each phase — waits at barriers included — to
// P.phase_cycles[warp * 16 + phase].
#ifdef JRLQP_PHASE_TIMING
#  define PH_DECL     \
    ph_t = clock64(); \
    for(int k_ = 0; k_ < 16; ++k_) ph_acc[k_] = 0
#  define PH_MARK(id)              \
    do                             \
    {                              \
      long long now_ = clock64();  \
      ph_acc[id] += now_ - ph_t;   \
      ph_t = now_;                 \
    } while(0)
#  define PH_FLUSH                                                                                      \
    do                                                                                                  \
    {                                                                                                   \
      if(lane == 0 && P.phase_cycles)                                                                   \
        for(int k_ = 0; k_ < 16; ++k_) atomicAdd(P.phase_cycles + warp * 16 + k_, (unsigned long long)ph_acc[k_]); \
    } while(0)
#else
#  define PH_DECL
#  define PH_MARK(id)
#  define PH_FLUSH
#endif
#define JRLQP_NONE 0x7fffffff

__device__ __forceinline__ double warp_sum32(double acc)
{
  // dot32 butterfly: acc[l] += acc[l ^ off], off = 16,8,4,2,1 (addition commutes => all lanes agree)
#pragma unroll
  for(int off = 16; off >= 1; off >>= 1) acc = acc + __shfl_xor_sync(JRLQP_FULL, acc, off);
  return acc;
}

template<int S>
__device__ __forceinline__ double pick(const double (&v)[S], int slot)
{
  double r = v[0];
#pragma unroll
  for(int s = 1; s < S; ++s)
    if(slot == s) r = v[s];
  return r;
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

// dot4 of two unit-stride vectors, evaluated redundantly by every calling thread (uniform result).
__device__ __noinline__ double dot4_uniform(int len, const double * __restrict__ a, const double * __restrict__ b)
{
  double c0 = 0, c1 = 0, c2 = 0, c3 = 0;
  int k = 0;
#pragma unroll 1
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

// dot4(ci, x) for one constraint normal read from global memory / L2 (thread = constraint: every lane
// walks its own row, so the loads of a chunk are issued together — CH values in flight per thread,
// as 128-bit loads when the rows are 16-byte aligned — instead of one round trip per 4 values).
template<bool VEC, int CH>
__device__ __forceinline__ double dot4_row(const double * __restrict__ ci, const double * xs, int n, const bool act = true)
{
  // act == false (constraint already active, or past the last one): the lane issues no memory request at all
  if(!act) ci = nullptr;
  double c0 = 0, c1 = 0, c2 = 0, c3 = 0;
  int k = 0;
#pragma unroll 1
  for(; k + CH - 1 < n; k += CH)
  {
    double v[CH];
    if(VEC)
    {
#pragma unroll
      for(int u = 0; u < CH / 2; ++u)
      {
        double2 t = make_double2(0.0, 0.0);
        if(ci != nullptr) t = *reinterpret_cast<const double2 *>(ci + k + 2 * u);
        v[2 * u] = t.x;
        v[2 * u + 1] = t.y;
      }
    }
    else
    {
#pragma unroll
      for(int u = 0; u < CH; ++u) v[u] = ci != nullptr ? ci[k + u] : 0.0;
    }
#pragma unroll
    for(int u = 0; u < CH / 4; ++u)
    {
#if JRLQP_OPT_V2_SCAN // measured: -0.5 % at n = 50, -1.2 % at n = 20 (profiles/r01zf_*): off; the d = J^T n+ loop keeps its 128-bit loads
      // xs is 16-byte aligned and k + 4 u even: two 128-bit broadcast loads instead of four 64-bit ones
      const double2 x01 = *reinterpret_cast<const double2 *>(xs + k + 4 * u);
      const double2 x23 = *reinterpret_cast<const double2 *>(xs + k + 4 * u + 2);
      c0 = fma(v[4 * u], x01.x, c0);
      c1 = fma(v[4 * u + 1], x01.y, c1);
      c2 = fma(v[4 * u + 2], x23.x, c2);
      c3 = fma(v[4 * u + 3], x23.y, c3);
#else
      c0 = fma(v[4 * u], xs[k + 4 * u], c0);
      c1 = fma(v[4 * u + 1], xs[k + 4 * u + 1], c1);
      c2 = fma(v[4 * u + 2], xs[k + 4 * u + 2], c2);
      c3 = fma(v[4 * u + 3], xs[k + 4 * u + 3], c3);
#endif
    }
  }
  {
    // remainder (< CH values): loads first, then the FMAs
    double v[CH];
    if(VEC)
    {
#pragma unroll
      for(int u = 0; u < CH / 2; ++u)
      {
        if(k + 2 * u + 1 < n && ci != nullptr)
        {
          const double2 t = *reinterpret_cast<const double2 *>(ci + k + 2 * u);
          v[2 * u] = t.x;
          v[2 * u + 1] = t.y;
        }
        else
        {
          v[2 * u] = (k + 2 * u < n && ci != nullptr) ? ci[k + 2 * u] : 0.0;
          v[2 * u + 1] = 0.0;
        }
      }
    }
    else
    {
#pragma unroll
      for(int u = 0; u < CH; ++u) v[u] = (k + u < n && ci != nullptr) ? ci[k + u] : 0.0;
    }
#pragma unroll
    for(int u = 0; u < CH / 4; ++u)
    {
      if(k + 4 * u < n) c0 = fma(v[4 * u], xs[k + 4 * u], c0);
      if(k + 4 * u + 1 < n) c1 = fma(v[4 * u + 1], xs[k + 4 * u + 1], c1);
      if(k + 4 * u + 2 < n) c2 = fma(v[4 * u + 2], xs[k + 4 * u + 2], c2);
      if(k + 4 * u + 3 < n) c3 = fma(v[4 * u + 3], xs[k + 4 * u + 3], c3);
    }
  }
  return (c0 + c1) + (c2 + c3);
}

// dot4(ci, x) for one constraint normal read from the TRANSPOSED copy of C (element k of constraint c at
// Ct[k * ld + c], global memory / L2): consecutive lanes = consecutive constraints read consecutive
// addresses, so a warp-level load is two full cache lines instead of 32 sectors of 32 different rows.
// L1 is bypassed (the slice is rewritten by this CTA for every problem).
// The transposed copy is stored in groups of JRLQP_CT_LD = 128 constraints, element k of constraint c at
// Ct[((c >> 7) * n + k) * 128 + (c & 127)]: the stride between consecutive k is a compile-time constant, so every
// load of a chunk is one instruction with an immediate offset (a run-time leading dimension cost a 64-bit multiply
// per load: 242 k of the 2.19 M warp-instructions of an n = 128 solve, profiles/r01zg_*).
#define JRLQP_CT_LD 128
__device__ __forceinline__ long long ct_offset(int c, int n) { return ((long long)(c >> 7) * n) * JRLQP_CT_LD + (c & (JRLQP_CT_LD - 1)); }
__host__ __device__ inline long long ct_doubles(int n, int mc) { return (long long)((mc + JRLQP_CT_LD - 1) / JRLQP_CT_LD) * n * JRLQP_CT_LD; }

// EVL: the loads carry an L2 evict_last policy (createpolicy + ld.global.cg.L2::cache_hint). For the large-n kernel, whose
// transposed C is ONE 5 MB copy shared by every CTA while 296 workspaces of 1.2 MB stream through L2: config C +7 %
// (profiles/r5l_*); the per-CTA copies of the four-warp kernel lose 3 % with it, hence a template parameter.
#ifndef JRLQP_CT_EVICT_LAST
#  define JRLQP_CT_EVICT_LAST 1
#endif
template<bool EVL>
__device__ __forceinline__ double ct_load(const double * p, const unsigned long long pol)
{
  if(EVL)
  {
    double v;
    asm volatile("ld.global.cg.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    return v;
  }
  return __ldcg(p);
}

template<int CH, bool EVL = false>
__device__ __forceinline__ double dot4_col(const double * ci, const double * xs, int n, const bool act = true)
{
  constexpr int ld = JRLQP_CT_LD;
  double c0 = 0, c1 = 0, c2 = 0, c3 = 0;
  int k = 0;
  unsigned long long pol = 0ull;
  if(EVL) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
#pragma unroll 1
  for(; k + CH - 1 < n; k += CH)
  {
    double v[CH];
#pragma unroll
    for(int u = 0; u < CH; ++u) v[u] = act ? ct_load<EVL>(ci + (k + u) * ld, pol) : 0.0;
#pragma unroll
    for(int u = 0; u < CH / 4; ++u)
    {
      c0 = fma(v[4 * u], xs[k + 4 * u], c0);
      c1 = fma(v[4 * u + 1], xs[k + 4 * u + 1], c1);
      c2 = fma(v[4 * u + 2], xs[k + 4 * u + 2], c2);
      c3 = fma(v[4 * u + 3], xs[k + 4 * u + 3], c3);
    }
  }
  {
    double v[CH];
#pragma unroll
    for(int u = 0; u < CH; ++u) v[u] = (k + u < n && act) ? ct_load<EVL>(ci + (k + u) * ld, pol) : 0.0;
#pragma unroll
    for(int u = 0; u < CH / 4; ++u)
    {
      if(k + 4 * u < n) c0 = fma(v[4 * u], xs[k + 4 * u], c0);
      if(k + 4 * u + 1 < n) c1 = fma(v[4 * u + 1], xs[k + 4 * u + 1], c1);
      if(k + 4 * u + 2 < n) c2 = fma(v[4 * u + 2], xs[k + 4 * u + 2], c2);
      if(k + 4 * u + 3 < n) c3 = fma(v[4 * u + 3], xs[k + 4 * u + 3], c3);
    }
  }
  return (c0 + c1) + (c2 + c3);
}

// Givens recurrence for the sweep i = n-2 .. q (the constraint would become the (q+1)-th).
// Serial part: rho_i = r(d[i], rho_{i+1}) — one division and one square root per link, exactly the
// operations of Eigen's makeGivens that feed r. The rest of makeGivens (second division, products)
// does not feed the recurrence: it is evaluated afterwards, one rotation per lane.
// Free function (one warp, lane = threadIdx.x & 31) shared by the shared-memory kernel and the
// global-workspace kernel (gi_large.cuh).
// (c, s) of rotation i from what the recurrence stashed: (t, u) and the branch taken in makeGivens
__device__ __forceinline__ double2 givens_cs(const int kind8, const double a, const double u)
{
  const int kind = kind8 & 7; // bit 3: "quotient proven after the loop" (givens_chain)
  double c, sn;
  if(kind == 0)
  {
    c = a < 0.0 ? -1.0 : 1.0;
    sn = 0.0;
  }
  else if(kind == 1)
  {
    c = 0.0;
    sn = a < 0.0 ? 1.0 : -1.0;
  }
  else
  {
    // kind 2: c = 1/u, s = -t c ; kind 3: s = -1/u, c = -t s  (-1/u == -(1/u) exactly)
    const double inv = 1.0 / u;
    if(kind == 2)
    {
      c = inv;
      sn = -a * c;
    }
    else
    {
      sn = -inv;
      c = -a * sn;
    }
  }
  return make_double2(c, sn);
}

// One link of the Givens recurrence outside the straight-line fast path (a zero operand, |d[i]| > |rho|, or the
// careful pass after a declined proof): Eigen's makeGivens branches, literally. Out of line (cold).
struct GivensLink
{
  double a, u, r;
  int kind;
};
__device__ __noinline__ GivensLink givens_link_slow(const double p, const double rho, const double rp, const bool have_rp)
{
  GivensLink g;
  if(rho == 0.0)
  {
    g.kind = 0;
    g.a = p;
    g.u = 1.0;
    g.r = fabs(p);
  }
  else if(p == 0.0)
  {
    g.kind = 1;
    g.a = rho;
    g.u = 1.0;
    g.r = fabs(rho);
  }
  else
  {
    const bool pg = fabs(p) > fabs(rho);
    g.kind = pg ? 2 : 3;
    const double num = pg ? rho : p, den = pg ? p : rho;
    bool ok = false;
    double a = 0.0;
    if(pg && have_rp) a = div_rcp(num, den, rp, ok); // 1 / p was precomputed
    if(!ok) a = num / den;
    const double uu = sqrt(fma(a, a, 1.0));
    g.a = a;
    g.u = den < 0.0 ? -uu : uu;
    g.r = den * g.u;
  }
  return g;
}

// ~ 1 / sqrt(x) to about 44 bits (hardware seed + one Newton step): all the seeds of the Givens recurrence need
__device__ __forceinline__ double rsqrt_seed(const double x)
{
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
  const double e = fma(-(x * y0), y0, 1.0);
  return fma(0.5 * y0, e, y0);
}

// SPLIT: the lane-parallel parts on either side of the recurrence (the reciprocals 1 / d[i] before it, the
// (c, s) pairs after it) are done by the caller with all the threads of the CTA, off this warp's critical path.
// gr: n doubles of scratch (the incoming rho of every link, for the deferred proofs).
//
// Latency of one link is what matters here. Per link, Eigen's makeGivens needs t = num / den, u = sqrt(1 + t^2),
// r = den u, with den = rho (the running value) in the common case |rho| >= |p|: a division and a square root that
// depend on the previous link — about 20 dependent FP64 operations with the stock sequences, 12 with a reciprocal
// carried along the recurrence (round 1). Here the link is cut to 7 dependent operations:
//  * SEEDS, computed for all links at once before the recurrence: rho_j^2 is (up to rounding) the suffix sum
//    S_j = d_j^2 + d_{j+1}^2 + ..., so rs_i = rsqrt(S_{i+1}) ~ 1 / rho_{i+1} and ys_i = rsqrt(1 + (d_i rs_i)^2) ~ 1 / u_i
//    are known to ~45 bits without running the recurrence (one warp scan + two rsqrt per lane);
//  * the link itself: t = p / rho by one correction of p rs_i (e = p - rho q0; t = q0 + e rs_i), s = 1 + t^2,
//    u = sqrt(s) by one correction of s ys_i (g = s ys; u = g + (s - g^2) ys / 2), r = |rho| u;
//  * PROOFS, evaluated for all links at once after the recurrence (one link per lane), that the quotient and the
//    square root so obtained are the correctly rounded ones — exact remainder tests: |p - rho t| < |rho| ulp(t) / 2
//    (see fp64_exact.cuh) and |s - u^2| < u ulp(u) (1 - 2^-41) (s - u^2 is exact for u within an ulp of the root, and
//    u = RN(sqrt s) iff (u - ulp/2)^2 < s < (u + ulp/2)^2). The results therefore never depend on the seeds: if one
//    proof is declined (rare: a seed off by more than ~2^-30, a result within 2^-90 of a rounding boundary) the
//    chain is redone with every quotient and root taken literally;
//  * every other case (|d[i]| > |rho|, a zero operand) leaves the straight-line fast path through one branch and
//    is evaluated with precomputed 1 / d[i] or literally.
template<bool SPLIT = false, int UNR = 2>
__device__ __forceinline__ void givens_chain(const int q, const int n, const int lane, const double * ds, double2 * gcs, double * gc, double * gs, int * gk, double * scr, double * gr)
{
  double * rpv = reinterpret_cast<double *>(gcs); // 1 / d[i]; gcs is only written after the chain
  if(!SPLIT)
  {
#pragma unroll 1
    for(int i = q + lane; i <= n - 1; i += 32)
    {
      rpv[i] = 1.0 / ds[i];
    }
  }
  // ---- seeds: element j = top - lane of d, chunks of 32 from the last element down; suffix sums by a warp scan.
  //      The seeds of link i = j - 1 go where the loop will later stash (t, u) of that link: gc[i], gs[i] (the loop
  //      reads the seeds of link i - 1 while it works on link i, and only then overwrites those of link i).
  {
    double carry = 0.0;
#pragma unroll 1
    for(int top = n - 1; top >= q; top -= 32)
    {
      const int j = top - lane;
      const bool valid = j >= q;
      const double dj = valid ? ds[j] : 0.0;
      const double dm = valid ? ds[j - 1] : 0.0; // d of the link this element feeds (j == 0: an unused read of the padding before d)
      double sc = dj * dj;
#pragma unroll
      for(int off = 1; off < 32; off <<= 1)
      {
        const double t = __shfl_up_sync(JRLQP_FULL, sc, off);
        if(lane >= off) sc = sc + t;
      }
      const double S = carry + sc;
      carry = __shfl_sync(JRLQP_FULL, S, 31);
      double rsd = rsqrt_seed(S);
      if(j == n - 1) rsd = dj < 0.0 ? -rsd : rsd; // the first rho is d[n-1] itself, sign included
      const double t = dm * rsd;
      const double ysd = rsqrt_seed(fma(t, t, 1.0));
      if(valid)
      {
        if(j > q)
        {
          gc[j - 1] = rsd;
          gs[j - 1] = ysd;
        }
        else
          scr[11] = rsd; // ~ 1 / rho_q: reciprocal of the new diagonal entry of R (an approximation is all rinv needs)
      }
    }
    __syncwarp();
  }
  bool careful = false;
#pragma unroll 1
  for(;;)
  {
    double rho = ds[n - 1];
    int i = n - 2;
    double p = ds[max(i, 0)];
    double rsi = gc[max(i, 0)], ysi = gs[max(i, 0)];
    double q0s = p * rsi; // first product of the quotient p / rho, issued ahead
#pragma unroll UNR
    for(; i >= q; --i)
    {
      // operands of the next link (i == 0: unused reads of the padding stored before the vectors)
      const double pn = ds[i - 1], rsn = gc[i - 1], ysn = gs[i - 1];
      // ---- fast path: t = p / rho, u = sqrt(1 + t^2), r = |rho| u, straight line
      const double e3 = fma(-rho, q0s, p);
      double a = fma(e3, rsi, q0s);
      const double s = fma(a, a, 1.0);
      const double g = s * ysi;
      const double rem = fma(-g, g, s);
      const double us = fma(rem, 0.5 * ysi, g);
      double r = fabs(rho) * us; // == rho * u with u = sign(rho) us, bit for bit
      double u = __hiloint2double(__double2hiint(us) | (__double2hiint(rho) & 0x80000000), __double2loint(us));
      int kind = 3 | 8; // bit 3: the quotient and the root of this link are to be proven after the loop
      const bool fast = !careful && fabs(p) <= fabs(rho) && p != 0.0;
      if(!fast)
      {
        // the trivial branches of makeGivens stay in line (structured problems — bound normals against a triangular J — feed
        // long runs of exact zeros: a call per link cost 13 % at n = 210, profiles/r03d_bench_C2_cold.json)
        if(rho == 0.0)
        {
          kind = 0;
          a = p;
          u = 1.0;
          r = fabs(p);
        }
        else if(p == 0.0)
        {
          kind = 1;
          a = rho;
          u = 1.0;
          r = fabs(rho);
        }
        else
        {
          const GivensLink g = givens_link_slow(p, rho, rpv[i], true);
          a = g.a;
          u = g.u;
          r = g.r;
          kind = g.kind;
        }
      }
      if(lane == 0)
      {
        // stash (t, u), the branch taken and the divisor of the quotient to be proven
        gc[i] = a;
        gs[i] = u;
        gk[i] = kind;
        gr[i] = rho;
      }
      rho = r;
      p = pn;
      rsi = rsn;
      ysi = ysn;
      q0s = pn * rsn;
    }
    if(lane == 0) scr[10] = rho;
    __syncwarp();
    if(careful) break;
    // ---- deferred proofs, one link per lane
    bool ok = true;
#pragma unroll 1
    for(int j = q + lane; j <= n - 2; j += 32)
    {
      if(gk[j] & 8)
      {
        const double a = gc[j], rh = gr[j], us = fabs(gs[j]);
        // t == RN(p / rho) iff |p - rho t| < |rho| ulp(t) / 2, t not a power of two, nothing near underflow
        const double e2 = fma(-rh, a, ds[j]); // exact remainder
        const int ahi = __double2hiint(a);
        const double hu = __hiloint2double((ahi & 0x7ff00000) - 0x03500000, 0);
        const double tol = fabs(rh) * hu;
        const bool pow2 = ((ahi & 0xfffff) | __double2loint(a)) == 0;
        // u == RN(sqrt(s)), s = 1 + t^2 in [1, 2] (so ulp(u) = 2^-52): |s - u^2| < u 2^-52 (1 - 2^-41)
        const double s = fma(a, a, 1.0);
        const double rem2 = fma(-us, us, s);
#ifndef JRLQP_DIAG_NOPROOF
        ok = ok && fabs(e2) < tol && tol > 1e-270 && !pow2 && fabs(rem2) < us * 0x1.ffffffffffp-53 && us >= 1.0 && us < 2.0;
#endif
      }
    }
    if(__all_sync(JRLQP_FULL, ok)) break;
    careful = true;
  }
  if(!SPLIT)
  {
#pragma unroll 1
    for(int i = q + lane; i <= n - 2; i += 32)
    {
      gcs[i] = givens_cs(gk[i], gc[i], gs[i]);
    }
  }
}

// r = R^-1 d(0:q) with the stock division in every link: the fallback of GiCta::back_substitution when one of its
// deferred proofs is declined (rare). Out of line: the instruction cache holds the fast loops only.
template<int W>
__device__ __noinline__ void back_substitution_exact(const int q, const int lane, const double * ds, const double * Rp, double * rs)
{
  double w[W], rr[W];
#pragma unroll
  for(int s = 0; s < W; ++s)
  {
    w[s] = lane + 32 * s < q ? ds[lane + 32 * s] : 0.0;
    rr[s] = 0.0;
  }
#pragma unroll 1
  for(int k = q - 1; k >= 0; --k)
  {
    const double * Rk = Rp + ((k * (k + 1)) >> 1);
    const double wk = __shfl_sync(JRLQP_FULL, pick<W>(w, k >> 5), k & 31);
    const double rk = wk / Rk[k];
#pragma unroll
    for(int s = 0; s < W; ++s)
    {
      const int r = lane + 32 * s;
      if(r == k)
        rr[s] = rk;
      else if(r < k)
        w[s] = fma(-rk, Rk[r], w[s]);
    }
  }
#pragma unroll
  for(int s = 0; s < W; ++s)
    if(lane + 32 * s < q) rs[lane + 32 * s] = rr[s];
}

struct Sel
{
  int p;
  int st;
};

// Exact sequential restatement of selectViolatedConstraint_ (src/GoldfarbIdnaniSolver.cpp:84-134),
// every thread redundantly. Only used when some constraint has BOTH slacks negative (bl > bu), the
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

template<int W, bool STAGE_C, bool WARM = false>
struct GiCta
{
  static constexpr int T = 32 * W;
#ifndef JRLQP_OPT_RS
#  define JRLQP_OPT_RS 1
#endif
#ifndef JRLQP_CHOL_LPR
#  define JRLQP_CHOL_LPR 0 // bit 0: one-warp kernel, bit 1: two-warp kernel — last columns of the Cholesky with 2 / 4 lanes per row.
                           // Measured and rejected (profiles/r5d_ab_*.txt: bit-identical, -1.8 % at n = 50, -17 % at n = 20): at 6 / 16 CTAs
                           // per SM the narrow kernels are bound by issue slots and the shared-memory pipe, not by the length of a
                           // column step; the strided 4-lane row reads add bank conflicts and both warps now execute every column.
#endif
#ifndef JRLQP_QR4
#  define JRLQP_QR4 1 // warm start: the Householder QR of the active normals updates its trailing columns with four lanes per column
#endif
#ifndef JRLQP_OPT_V2
#  define JRLQP_OPT_V2 1
#endif
#ifndef JRLQP_OPT_PRED
#  define JRLQP_OPT_PRED 1
#endif
#ifndef JRLQP_OPT_CS
#  define JRLQP_OPT_CS 1
#endif
#ifndef JRLQP_OPT_T1
#  define JRLQP_OPT_T1 1
#endif
#ifndef JRLQP_OPT_ADD
#  define JRLQP_OPT_ADD 1
#endif
#ifndef JRLQP_CH1
#  define JRLQP_CH1 4
#endif
#ifndef JRLQP_CH2
#  define JRLQP_CH2 8
#endif
#ifndef JRLQP_CH4
#  define JRLQP_CH4 16
#endif
  static constexpr int CH = W == 1 ? JRLQP_CH1 : (W == 2 ? JRLQP_CH2 : JRLQP_CH4); // values in flight per thread in the constraint scan
#ifndef JRLQP_PF1
#  define JRLQP_PF1 2
#endif
#ifndef JRLQP_PF4
#  define JRLQP_PF4 4
#endif
#ifndef JRLQP_UNR_DZ
#  define JRLQP_UNR_DZ 2
#endif
#ifndef JRLQP_PF2
#  define JRLQP_PF2 2
#endif
  // inner loops of the Cholesky and of J = L^-T: unrolled four times in the wide kernels (+3 % at n = 128; -0.5 % at n = 50 and
  // -9 % at n = 20, profiles/r01zn_ab_*.txt)
#ifndef JRLQP_UNR_CHOL_NARROW
#  define JRLQP_UNR_CHOL_NARROW 1
#endif
#ifndef JRLQP_UNR_JB_NARROW
#  define JRLQP_UNR_JB_NARROW 1
#endif
  static constexpr int UNR_CHOL = W >= 3 ? JRLQP_UNR_CHOL : JRLQP_UNR_CHOL_NARROW, UNR_JB = W >= 3 ? JRLQP_UNR_JB : JRLQP_UNR_JB_NARROW;
  static constexpr int UNR_DZ = W >= 3 ? 4 : JRLQP_UNR_DZ; // unroll factor of the d = J^T n+ and z = J2 d2 loops
  static constexpr int PF = W == 1 ? JRLQP_PF1 : (W == 2 ? JRLQP_PF2 : JRLQP_PF4); // rotations per chunk when the Givens table is applied
  // ---- immutable per-launch
  const GiParams & P;
  const int tid, lane, warp;
  const int n, mc, nb, m, ldj;
  double *Jb, *Rp, *xs, *zs, *ds, *rs, *us, *cv, *ldiag, *rinv, *Cs, *scr;
  double2 * grec; // per-link records of the Givens sweep: (t, u) then (c, s) ; (rho, branch) — see givens_recurrence
  int * alist;
  int * iscr;
  unsigned short * ilist; // compacted list of the inactive general constraints (constraint scan)
  signed char * stat;
  double *Vp, *bact, *hco, *alp; // warm start: Householder vectors (packed), b_act, tau, alpha = J^T a
  signed char * eqf; // 1 where bl == bu (resp. xl == xu): constraints initActiveSet pre-activates
  // ---- per-problem views
  const double *Cb, *bl, *bu, *xl, *xu;
  long long ldC; // leading dimension of the constraint-normal storage Cb (staged or global)
  // The transposed-copy scan exists in the wide kernels only (n > 64, where it measures faster): in the narrow ones its
  // code would sit, never executed, inside the hottest loop of an instruction-cache-sensitive kernel.
  static constexpr bool CT_SCAN = !STAGE_C && (W >= 3 || JRLQP_CT_ALLW);
  bool cvec; // rows of Cb are 16-byte aligned (128-bit loads allowed)
  double * Ct = nullptr; // this CTA's slice for the transposed copy of C (null: scan C in place)
  bool ct_valid = false; // the slice holds the (batch-shared) C already
  // ---- solver state (uniform across threads)
  int q;
  double f;
  bool cv_ready = false; // cv already holds the normal of the next step's constraint (loaded while the rotations were applied)
#ifdef JRLQP_PHASE_TIMING
  long long ph_t, ph_acc[16];
#endif

  __device__ GiCta(const GiParams & p, double * smem)
  : P(p), tid(threadIdx.x), lane(threadIdx.x & 31), warp(threadIdx.x >> 5), n(p.n), mc(p.mc), nb(p.nb), m(p.mc + p.nb), ldj(p.ldj)
  {
    Jb = smem;
    Rp = smem + p.off_R;
    xs = smem + p.off_x;
    zs = smem + p.off_z;
    ds = smem + p.off_d;
    rs = smem + p.off_r;
    us = smem + p.off_u;
    cv = smem + p.off_cv;
    grec = reinterpret_cast<double2 *>(smem + p.off_gcs);
    ldiag = smem + p.off_ldiag;
    rinv = smem + p.off_rinv;
    scr = smem + p.off_scr;
    Cs = smem + p.off_C;
    alist = reinterpret_cast<int *>(smem + p.off_alist);
    iscr = reinterpret_cast<int *>(smem + p.off_iscr);
    stat = reinterpret_cast<signed char *>(smem + p.off_stat);
    eqf = reinterpret_cast<signed char *>(smem + p.off_eq);
    ilist = reinterpret_cast<unsigned short *>(smem + p.off_il);
    Vp = smem + p.off_V;
    bact = smem + p.off_bact;
    hco = smem + p.off_hco;
    alp = smem + p.off_alpha;
  }

  __device__ __forceinline__ void sync() const
  {
    if(W == 1)
      __syncwarp();
    else
      __syncthreads();
  }
  __device__ __forceinline__ static int colR(int k) { return (k * (k + 1)) >> 1; }

  // Transposed copy of this problem's C into the CTA's global slice (it stays in L2 and is read by every
  // constraint scan of the solve). C is read once, coalesced (thread = k along a normal), passed through
  // the J buffer (free until init) as an [constraint][k] tile with odd leading dimension, and written
  // coalesced (thread = constraint). Called before init() / init_warm().
  __device__ void stage_ct(long long b)
  {
    if(!CT_SCAN || Ct == nullptr || mc == 0) return;
    if(P.sC == 0 && ct_valid) return; // C shared by the batch: transposed once per CTA
    const double * __restrict__ Cg = P.C + b * P.sC;
    const int ldt = n | 1;
    const int chunk = max(1, (n * ldj) / ldt);
#pragma unroll 1
    for(int c0 = 0; c0 < mc; c0 += chunk)
    {
      const int cn = min(chunk, mc - c0);
      if(tid < n)
      {
#pragma unroll 4
        for(int c = 0; c < cn; ++c) Jb[c * ldt + tid] = __ldg(Cg + (long long)(c0 + c) * P.ldc + tid);
      }
      sync();
#pragma unroll 1
      for(int cc = tid; cc < cn; cc += T)
      {
        const double * src = Jb + cc * ldt;
        double * dst = Ct + ct_offset(c0 + cc, n);
#pragma unroll 4
        for(int k = 0; k < n; ++k) __stcg(dst + k * JRLQP_CT_LD, src[k]);
      }
      sync();
    }
    ct_valid = true;
  }

  // ------------------------------------------------------------------------------------------
  // init_ (src/GoldfarbIdnaniSolver.cpp:56-82): Cholesky, J = L^-T, x = -G^-1 a, f = a.x/2
  // ------------------------------------------------------------------------------------------
  // One column of the left-looking Cholesky with G = 2 or 4 lanes per row, for the last T / G columns (rows n - T / G ... n - 1
  // <-> lane groups 0 ... T / G - 1): lane c of a group runs the chains c, c + G, ... of the canonical dot4 over ITS entries
  // j = c (mod G) of the row (G = 4: chain c; G = 2: chains 2 c and 2 c + 1), the chains are folded (a0 + a1) + (a2 + a3) by
  // shuffles (additions commute), lane 0 of the group forms and stores the entry. Same operations in the same order for every
  // entry as the one-lane-per-row step in init(): same bits; the long inner products of the last columns cost a half / a quarter
  // and both warps share them. Returns false on a non-positive pivot (uniform).
  template<int G, bool WRITE_RS = false>
  __device__ __forceinline__ bool chol_step_split(const int k)
  {
    const int g = tid / G, c = tid % G;
    const int i = n - T / G + g; // (< k: this group has no row in column k)
    const bool act = i >= k;
    const double * Li = Jb + (act ? i : k) * ldj;
    const double * Lk = Jb + k * ldj;
    double sum;
    if(G == 4)
    {
      double a = 0.0;
#pragma unroll 2
      for(int j = c; j < k; j += 4) a = fma(Li[j], Lk[j], a);
      a = a + __shfl_xor_sync(JRLQP_FULL, a, 1);
      sum = a + __shfl_xor_sync(JRLQP_FULL, a, 2);
    }
    else
    {
      double a0 = 0.0, a1 = 0.0;
      int j = 2 * c;
#pragma unroll 2
      for(; j + 1 < k; j += 4)
      {
        a0 = fma(Li[j], Lk[j], a0);
        a1 = fma(Li[j + 1], Lk[j + 1], a1);
      }
      if(j < k) a0 = fma(Li[j], Lk[j], a0);
      const double h = a0 + a1;
      sum = h + __shfl_xor_sync(JRLQP_FULL, h, 1);
    }
    const double v = Li[k] - sum;
    if(act && i == k && c == 0) scr[0] = v;
    sync();
    const double vk = scr[0];
    if(vk <= 0.0) return false; // Eigen llt: "if (x <= 0) return k" -> NON_POS_HESSIAN (uniform)
    const double lkk = sqrt(vk);
    if(act && c == 0)
    {
      if(i == k)
      {
        Jb[i * ldj + k] = lkk;
        ldiag[k] = lkk;
        if(WRITE_RS || !JRLQP_OPT_RS) rs[k] = 1.0 / lkk;
      }
      else
        Jb[i * ldj + k] = v / lkk;
    }
    sync();
    return true;
  }

  __device__ bool init(long long b)
  {
    const double * __restrict__ Gb = P.G + b * P.sG;
    const double * __restrict__ ab = P.a + b * P.sa;
    const int ldg = P.ldg;
    const int i = tid; // this thread's row (and, later, column)
    const int ic = min(i, n - 1);

    // stage the lower triangle of G (column-major in HBM: threads run down a column => coalesced)
#pragma unroll 4
    for(int j = 0; j < n; ++j)
      if(i < n && i >= j) Jb[i * ldj + j] = __ldg(Gb + i + (long long)j * ldg);
    if(STAGE_C)
    {
      // C is n x mc column-major: column c (one normal) is contiguous => coalesced along k
      const double * __restrict__ Cg = P.C + b * P.sC;
#pragma unroll 4
      for(int c = 0; c < mc; ++c)
        if(i < n) Cs[c * P.ldcs + i] = __ldg(Cg + i + (long long)c * P.ldc);
    }
    else if(!JRLQP_BULK_PREFETCH && mc > 0 && Ct == nullptr)
    {
      // pull this problem's C towards L2 while the factorisation runs (it is first needed by the scan)
      const char * Cg = reinterpret_cast<const char *>(P.C + b * P.sC);
      const long long bytes = ((long long)(mc - 1) * P.ldc + n) * 8;
      for(long long o = (long long)tid * 128; o < bytes; o += (long long)T * 128)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(Cg + o));
    }
    sync();

    // --- left-looking Cholesky, thread = row
    {
      const double * Li = Jb + ic * ldj;
      // DMMA (W = 4): at the first column of every 8-column panel the four dot4 accumulators of all the entries of the panel
      // are advanced over the columns j < K0 = 16 floor(k / 16) by mma.sync.m8n8k4.f64 — one DMMA is the sequential chain
      // fma(a3,b3, fma(a2,b2, fma(a1,b1, fma(a0,b0,c)))) bit for bit (profiles/r02t_dmma_probe.txt), so an accumulator tile fed
      // with j = jb + t, jb + t + 4, jb + t + 8, jb + t + 12 IS chain t of the canonical dot4 — and parked in the (still
      // unused) storage of R; the column loop below then starts its chains from them at j = K0 (<= 15 scalar terms per
      // entry instead of up to n - 1). Rows of a warp are produced and consumed by that warp: no block barrier.
      constexpr bool DMMA = W == 4 && JRLQP_DMMA_CHOL != 0;
      constexpr int ACS = 33; // doubles per row of the parked accumulators: [column of the panel][chain] + 1 (odd: conflict-free)
      double * const acs = Rp;
      const bool dmma_on = DMMA && (long long)(32 * W) * ACS <= (long long)n * (n + 1) / 2;
      // narrow kernels: once no more than T / 2 (T / 4) rows are left below the pivot, a row is given to TWO (FOUR) lanes
      // (chol_step_split) — the columns with the longest inner products are the ones with the fewest rows
      constexpr bool LPR = W <= 2 && ((JRLQP_CHOL_LPR >> (W - 1)) & 1) != 0;
      const int k1 = LPR ? max(0, n - T / 2) : n;
#pragma unroll 1
      for(int k = 0; k < k1; ++k)
      {
        int K0 = 0;
        if(DMMA && dmma_on)
        {
          K0 = (k >> 4) << 4;
          if((k & 7) == 0 && K0 > 0 && 32 * warp + 31 >= k)
          {
            const int g = lane >> 2, t = lane & 3; // fragments: A(row g, k t), B(k t, column g), C(row g, columns 2 t, 2 t + 1)
            double acc[4][4][2];
#pragma unroll
            for(int rt = 0; rt < 4; ++rt)
#pragma unroll
              for(int ch = 0; ch < 4; ++ch) acc[rt][ch][0] = acc[rt][ch][1] = 0.0;
            const double * Bp = Jb + min(k + g, n - 1) * ldj + 4 * t;
            const double * Ap = Jb + min(32 * warp + g, n - 1) * ldj + 4 * t;
            const int rstep = 8 * ldj;
            const bool r1 = 32 * warp + 8 + g < n, r2 = 32 * warp + 16 + g < n, r3 = 32 * warp + 24 + g < n; // (rows past n - 1: row of tile 0 again)
#pragma unroll 1
            for(int jb = 0; jb < K0; jb += 16)
            {
#pragma unroll
              for(int ch = 0; ch < 4; ++ch)
              {
                const double bf = Bp[jb + ch];
                const double a0f = Ap[jb + ch];
                const double a1f = Ap[(r1 ? rstep : 0) + jb + ch];
                const double a2f = Ap[(r2 ? 2 * rstep : 0) + jb + ch];
                const double a3f = Ap[(r3 ? 3 * rstep : 0) + jb + ch];
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(acc[0][ch][0]), "+d"(acc[0][ch][1]) : "d"(a0f), "d"(bf));
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(acc[1][ch][0]), "+d"(acc[1][ch][1]) : "d"(a1f), "d"(bf));
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(acc[2][ch][0]), "+d"(acc[2][ch][1]) : "d"(a2f), "d"(bf));
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(acc[3][ch][0]), "+d"(acc[3][ch][1]) : "d"(a3f), "d"(bf));
              }
            }
#pragma unroll
            for(int rt = 0; rt < 4; ++rt)
            {
              double * o = acs + (32 * warp + 8 * rt + g) * ACS + 8 * t;
#pragma unroll
              for(int ch = 0; ch < 4; ++ch)
              {
                o[ch] = acc[rt][ch][0]; // column 2 t of the panel
                o[4 + ch] = acc[rt][ch][1]; // column 2 t + 1
              }
            }
            __syncwarp();
          }
        }
        double v = 0.0;
        if(32 * warp + 31 >= k) // warps entirely above the pivot row have nothing to do
        {
          const double * Lk = Jb + k * ldj;
          double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
          int j = 0;
          if(DMMA && K0 > 0)
          {
            const double * o = acs + tid * ACS + 4 * (k & 7);
            a0 = o[0];
            a1 = o[1];
            a2 = o[2];
            a3 = o[3];
            j = K0;
          }
#pragma unroll UNR_CHOL
          for(; j + 3 < k; j += 4)
          {
            a0 = fma(Li[j], Lk[j], a0);
            a1 = fma(Li[j + 1], Lk[j + 1], a1);
            a2 = fma(Li[j + 2], Lk[j + 2], a2);
            a3 = fma(Li[j + 3], Lk[j + 3], a3);
          }
          if(j < k) a0 = fma(Li[j], Lk[j], a0);
          if(j + 1 < k) a1 = fma(Li[j + 1], Lk[j + 1], a1);
          if(j + 2 < k) a2 = fma(Li[j + 2], Lk[j + 2], a2);
          v = Li[k] - ((a0 + a1) + (a2 + a3));
          if(i == k) scr[0] = v;
        }
        sync();
        double vk = scr[0];
        if(vk <= 0.0) return false; // Eigen llt: "if (x <= 0) return k" -> NON_POS_HESSIAN (uniform)
        double lkk = sqrt(vk);
        if(i == k)
        {
          Jb[i * ldj + k] = lkk;
          ldiag[k] = lkk;
#if !JRLQP_OPT_RS
          rs[k] = 1.0 / lkk; // reciprocal of the diagonal for J = L^-T
#endif
        }
        else if(i > k && i < n)
          Jb[i * ldj + k] = v / lkk;
        sync();
      }
      if(LPR)
      {
        const int k2 = max(k1, n - T / 4);
#pragma unroll 1
        for(int k = k1; k < k2; ++k)
          if(!chol_step_split<2>(k)) return false;
#pragma unroll 1
        for(int k = k2; k < n; ++k)
          if(!chol_step_split<4>(k)) return false;
      }
#if JRLQP_OPT_RS
      // reciprocals of the diagonal (for J = L^-T and the two triangular solves): one division per thread,
      // after the loop, instead of one on the critical path of every column
      if(i < n) rs[i] = 1.0 / ldiag[i];
      sync();
#endif
    }

    // optional copy-out of L (what the reference leaves in G)
    if(P.L != nullptr)
    {
      double * Lout = P.L + b * (long long)n * n;
      for(int j = 0; j < n; ++j)
        if(i < n && i >= j) Lout[i + (long long)j * n] = Jb[i * ldj + j];
    }

    // --- x = -G^-1 a on warp 0 (serial division chain, rows over the lanes, W slots), while the
    //     other warps already build their columns of J = L^-T; warp 0 joins afterwards.
    if(warp == 0)
    {
      double y[W];
#pragma unroll
      for(int s = 0; s < W; ++s) y[s] = lane + 32 * s < n ? __ldg(ab + lane + 32 * s) : 0.0;
#pragma unroll 1
      for(int k = 0; k < n; ++k)
      {
        const double lk = ldiag[k], vk = __shfl_sync(JRLQP_FULL, pick<W>(y, k >> 5), k & 31);
        bool okd;
        double yk = div_rcp(vk, lk, rs[k], okd); // rs[k] = 1 / L(k,k)
        if(!okd) yk = vk / lk;
#pragma unroll
        for(int s = 0; s < W; ++s)
        {
          int r = lane + 32 * s;
          if(r == k)
            y[s] = yk;
          else if(r > k && r < n)
            y[s] = fma(-yk, Jb[r * ldj + k], y[s]);
        }
      }
#pragma unroll 1
      for(int k = n - 1; k >= 0; --k)
      {
        const double lk = ldiag[k], vk = __shfl_sync(JRLQP_FULL, pick<W>(y, k >> 5), k & 31);
        bool okd;
        double xk = div_rcp(vk, lk, rs[k], okd);
        if(!okd) xk = vk / lk;
#pragma unroll
        for(int s = 0; s < W; ++s)
        {
          int r = lane + 32 * s;
          if(r == k)
            y[s] = xk;
          else if(r < k)
            y[s] = fma(-xk, Jb[k * ldj + r], y[s]);
        }
      }
      double facc = 0.0;
#pragma unroll
      for(int s = 0; s < W; ++s)
      {
        int r = lane + 32 * s;
        if(r < n)
        {
          double xi = -y[s];
          xs[r] = xi;
          facc = fma(__ldg(ab + r), xi, facc);
        }
      }
      scr[1] = 0.5 * warp_sum32(facc);
    }

    // --- J = L^-T in place (upper triangle), thread = column j; a column only depends on L and on
    //     itself, so no synchronisation is needed between threads while it is built.
    {
      const int j = i;
      const int jc = ic;
      const int jmax = min(n - 1, 32 * warp + 31); // last column handled by this warp
      if(j < n) Jb[j * ldj + j] = rs[j];
#pragma unroll 1
      for(int r = jmax - 1; r >= 0; --r)
      {
        double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll UNR_JB
        for(int k0 = r + 1; k0 <= jmax; k0 += 4)
        {
          // accumulator index (k - r - 1) & 3. Loads are unconditional (rows past j, even past n - 1,
          // are addressable shared memory) and the accumulation is a select: no divergent branch.
          const double * Jk = Jb + k0 * ldj;
          const double t0 = fma(Jk[r], Jk[jc], a0);
          const double t1 = fma(Jk[ldj + r], Jk[ldj + jc], a1);
          const double t2 = fma(Jk[2 * ldj + r], Jk[2 * ldj + jc], a2);
          const double t3 = fma(Jk[3 * ldj + r], Jk[3 * ldj + jc], a3);
          a0 = k0 <= j ? t0 : a0;
          a1 = k0 + 1 <= j ? t1 : a1;
          a2 = k0 + 2 <= j ? t2 : a2;
          a3 = k0 + 3 <= j ? t3 : a3;
        }
        if(r < j && j < n) Jb[r * ldj + j] = (-((a0 + a1) + (a2 + a3))) * rs[r];
      }
    }
    sync();
    f = scr[1];
    // clear the strict lower triangle (L is no longer needed), thread = row
    if(i < n)
    {
#pragma unroll 4
      for(int c = 0; c < i; ++c) Jb[i * ldj + c] = 0.0;
    }
    // A_.reset()
    for(int c = tid; c < m; c += T)
    {
      stat[c] = ST_INACTIVE;
      eqf[c] = (c < mc ? (bl[c] == bu[c]) : (xl[c - mc] == xu[c - mc])) ? 1 : 0; // initActiveSet's tests, once
    }
    q = 0;
    sync();
    return true;
  }

  // ==========================================================================================
  // Warm-start capable initialisation: experimental::GoldfarbIdnaniSolver::init_
  // (src/experimental/GoldfarbIdnaniSolver.cpp:66-111) and the functions it calls. The active
  // normals N are reduced by B = L^-1 N, B = Q R (Householder, in place: R in the packed upper
  // storage Rp, the essential parts of the Householder vectors in the packed strictly-lower storage
  // Vp), J = L^-T Q; then the primal / dual point of the guessed active set. Arithmetic: the
  // canonical order of oracle/warm_oracle.cpp, bit for bit.
  // ==========================================================================================
  __device__ __forceinline__ int offV(int k) const { return k * (n - 1) - ((k * (k - 1)) >> 1); } // column k of Vp: rows k+1 .. n-1

  // element (i, k) of the n x q working matrix (B, then R over essential parts)
  __device__ __forceinline__ double * Bat(int i, int k) const { return i <= k ? Rp + colR(k) + i : Vp + offV(k) + (i - k - 1); }

  // processInitialActiveSet (src/experimental/GoldfarbIdnaniSolver.cpp:306-381). Returns the status.
  __device__ int warm_active_set(long long b)
  {
    const bool use_as = P.as_in != nullptr && P.warm_start != 0;
    const signed char * as = use_as ? P.as_in + b * P.s_as : nullptr;
    const double big = P.big_bnd;
    for(int c = tid; c < m; c += T)
    {
      int s = ST_INACTIVE;
      if(c >= mc)
      {
        const int i = c - mc;
        const double lo = xl[i], up = xu[i];
        if(lo == up)
          s = ST_FIXED;
        else if(use_as)
        {
          const int g = as[c];
          // FIXED is ignored (the bounds differ), so is a guess on an infinite bound; statuses that
          // do not describe a bound are ignored too (the reference asserts on them)
          if((g == ST_LOWER_BOUND && !(lo < -big)) || (g == ST_UPPER_BOUND && !(up > big))) s = g;
        }
      }
      else
      {
        const double lo = bl[c], up = bu[c];
        if(lo == up)
          s = ST_EQUALITY;
        else if(use_as)
        {
          const int g = as[c];
          if((g == ST_LOWER && !(lo < -big)) || (g == ST_UPPER && !(up > big)) || g == ST_EQUALITY) s = g;
        }
      }
      stat[c] = (signed char)s;
    }
    sync();
    // ordered active list: bounds first, then general constraints (activation order of the reference)
    if(warp == 0)
    {
      int cnt = 0, neq = 0;
      for(int base = 0; base < m; base += 32)
      {
        const int o = base + lane; // position in activation order
        const int c = o < nb ? mc + o : o - nb;
        const int sv = o < m ? stat[c] : ST_INACTIVE;
        const unsigned act = __ballot_sync(JRLQP_FULL, sv != ST_INACTIVE);
        neq += __popc(__ballot_sync(JRLQP_FULL, sv == ST_EQUALITY || sv == ST_FIXED));
        const int pos = cnt + __popc(act & ((1u << lane) - 1u));
        // more than n guesses: positions >= n are kept in a spill that only the trimming below reads;
        // alist has n entries, so the rare overflow case is resolved serially afterwards
        if(sv != ST_INACTIVE && pos < n) alist[pos] = c;
        cnt += __popc(act);
      }
      if(lane == 0)
      {
        iscr[0] = cnt;
        iscr[1] = neq;
      }
    }
    sync();
    int cnt = iscr[0];
    const int neq = iscr[1];
    if(cnt > n)
    {
      if(neq > n) return TS_OVERCONSTRAINED_PROBLEM;
      // "Work backward to deactivate inequality constraints until the number of constraints is nbVar":
      // walking the activation order backwards, every non-equality entry is dropped until n remain.
      if(tid == 0)
      {
        int excess = cnt - n;
        for(int o = m - 1; o >= 0 && excess > 0; --o)
        {
          const int c = o < nb ? mc + o : o - nb;
          const int sv = stat[c];
          if(sv != ST_INACTIVE && sv != ST_EQUALITY && sv != ST_FIXED)
          {
            stat[c] = ST_INACTIVE;
            --excess;
          }
        }
        int pos = 0;
        for(int o = 0; o < m; ++o)
        {
          const int c = o < nb ? mc + o : o - nb;
          if(stat[c] != ST_INACTIVE) alist[pos++] = c;
        }
      }
      cnt = n;
      sync();
    }
    q = cnt;
    return TS_SUCCESS;
  }

  // initializePrimalDualPoints (src/experimental/GoldfarbIdnaniSolver.cpp:461-486)
  __device__ void warm_primal_dual(const double * ab)
  {
    const int j = tid, jc = min(tid, n - 1);
    // alpha = J^T a, thread = column (cv holds a)
    if(j < n) cv[j] = __ldg(ab + j);
    sync();
    {
      double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
      const double * Jc = Jb + jc;
      int i = 0;
      for(; i + 3 < n; i += 4)
      {
        a0 = fma(Jc[i * ldj], cv[i], a0);
        a1 = fma(Jc[(i + 1) * ldj], cv[i + 1], a1);
        a2 = fma(Jc[(i + 2) * ldj], cv[i + 2], a2);
        a3 = fma(Jc[(i + 3) * ldj], cv[i + 3], a3);
      }
      if(i < n) a0 = fma(Jc[i * ldj], cv[i], a0);
      if(i + 1 < n) a1 = fma(Jc[(i + 1) * ldj], cv[i + 1], a1);
      if(i + 2 < n) a2 = fma(Jc[(i + 2) * ldj], cv[i + 2], a2);
      if(j < n) alp[j] = (a0 + a1) + (a2 + a3);
    }
    // beta = R^-T b_act on warp 0 (column-oriented forward substitution, true division), into zs
    if(warp == 0)
    {
      double w[W];
#pragma unroll
      for(int s = 0; s < W; ++s) w[s] = lane + 32 * s < q ? bact[lane + 32 * s] : 0.0;
      for(int k = 0; k < q; ++k)
      {
        const double bk = __shfl_sync(JRLQP_FULL, pick<W>(w, k >> 5), k & 31) / Rp[colR(k) + k];
#pragma unroll
        for(int s = 0; s < W; ++s)
        {
          const int i = lane + 32 * s;
          if(i == k)
            w[s] = bk;
          else if(i > k && i < q)
            w[s] = fma(-bk, Rp[colR(i) + k], w[s]);
        }
      }
#pragma unroll
      for(int s = 0; s < W; ++s)
        if(lane + 32 * s < q) zs[lane + 32 * s] = w[s];
    }
    sync();
    // x = J1 beta - J2 alpha2, thread = row ; d = alpha1 + beta (right-hand side of u)
    {
      const double * Jr = Jb + jc * ldj;
      double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
      int c = 0;
      for(; c + 3 < q; c += 4)
      {
        a0 = fma(Jr[c], zs[c], a0);
        a1 = fma(Jr[c + 1], zs[c + 1], a1);
        a2 = fma(Jr[c + 2], zs[c + 2], a2);
        a3 = fma(Jr[c + 3], zs[c + 3], a3);
      }
      if(c < q) a0 = fma(Jr[c], zs[c], a0);
      if(c + 1 < q) a1 = fma(Jr[c + 1], zs[c + 1], a1);
      if(c + 2 < q) a2 = fma(Jr[c + 2], zs[c + 2], a2);
      const double s1 = (a0 + a1) + (a2 + a3);
      a0 = a1 = a2 = a3 = 0;
      c = q;
      for(; c + 3 < n; c += 4)
      {
        a0 = fma(Jr[c], alp[c], a0);
        a1 = fma(Jr[c + 1], alp[c + 1], a1);
        a2 = fma(Jr[c + 2], alp[c + 2], a2);
        a3 = fma(Jr[c + 3], alp[c + 3], a3);
      }
      if(c < n) a0 = fma(Jr[c], alp[c], a0);
      if(c + 1 < n) a1 = fma(Jr[c + 1], alp[c + 1], a1);
      if(c + 2 < n) a2 = fma(Jr[c + 2], alp[c + 2], a2);
      const double s2 = (a0 + a1) + (a2 + a3);
      if(j < n) xs[j] = s1 - s2;
      if(j < q) ds[j] = alp[j] + zs[j];
    }
    sync();
    // u = R^-1 (alpha1 + beta)
    if(warp == 0)
    {
      back_substitution();
      __syncwarp();
#pragma unroll
      for(int s = 0; s < W; ++s)
        if(lane + 32 * s < q) us[lane + 32 * s] = rs[lane + 32 * s];
    }
    // f = beta.(0.5 beta + alpha1) - 0.5 |alpha2|^2 (dot32 order), every warp redundantly
    {
      double s1 = 0.0, s2 = 0.0;
      for(int k = lane; k < q; k += 32)
      {
        const double bk = zs[k];
        s1 = fma(bk, fma(0.5, bk, alp[k]), s1);
      }
      for(int k = lane; k < n - q; k += 32)
      {
        const double ak = alp[q + k];
        s2 = fma(ak, ak, s2);
      }
      f = warp_sum32(s1) - 0.5 * warp_sum32(s2);
    }
    sync();
  }

  // experimental init_: returns the termination status (SUCCESS: ready for the main loop)
  __device__ int init_warm(long long b, int & it)
  {
    const double * __restrict__ Gb = P.G + b * P.sG;
    const double * __restrict__ ab = P.a + b * P.sa;
    const int ldg = P.ldg;
    const int i = tid;
    const int ic = min(i, n - 1);

    // ---- sequences: the factor of an earlier step (same G, hence the same bits) comes back from HBM. The slot of diag(L)[0]
    //      doubles as the state of the entry: > 0 (or NaN) a factor, < 0 G is not positive definite, 0 nothing stored yet
    double * const fc = P.fcache_mode != 0 ? P.fcache + b * P.fcache_stride : nullptr;
    const double fstate = P.fcache_mode == 2 ? fc[(long long)n * n] : 0.0; // (uniform)

    int st = warm_active_set(b);
    if(st != TS_SUCCESS)
    {
      if(P.fcache_mode == 1 && tid == 0) fc[(long long)n * n] = 0.0; // this step stores nothing: a later one factorises
      return st;
    }

    const bool fload = fstate != 0.0;
    if(fload)
    {
      if(fstate < 0.0) return TS_NON_POS_HESSIAN;
#pragma unroll 4
      for(int j = 0; j < n; ++j)
        if(i < n) Jb[i * ldj + j] = fc[i + (long long)j * n];
      if(i < n)
      {
        ldiag[i] = fc[(long long)n * n + i];
        rs[i] = fc[(long long)n * n + n + i];
      }
      sync();
    }
    if(!fload)
    {
    // ---- Cholesky (same code path and order as init())
#pragma unroll 4
    for(int j = 0; j < n; ++j)
      if(i < n && i >= j) Jb[i * ldj + j] = __ldg(Gb + i + (long long)j * ldg);
    sync();
    {
      const double * Li = Jb + ic * ldj;
      constexpr bool LPR = W <= 2 && ((JRLQP_CHOL_LPR >> (W - 1)) & 1) != 0; // (as in init())
      const int k1 = LPR ? max(0, n - T / 2) : n;
#pragma unroll 1
      for(int k = 0; k < k1; ++k)
      {
        double v = 0.0;
        if(32 * warp + 31 >= k)
        {
          const double * Lk = Jb + k * ldj;
          double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
          int j = 0;
#pragma unroll 1
          for(; j + 3 < k; j += 4)
          {
            a0 = fma(Li[j], Lk[j], a0);
            a1 = fma(Li[j + 1], Lk[j + 1], a1);
            a2 = fma(Li[j + 2], Lk[j + 2], a2);
            a3 = fma(Li[j + 3], Lk[j + 3], a3);
          }
          if(j < k) a0 = fma(Li[j], Lk[j], a0);
          if(j + 1 < k) a1 = fma(Li[j + 1], Lk[j + 1], a1);
          if(j + 2 < k) a2 = fma(Li[j + 2], Lk[j + 2], a2);
          v = Li[k] - ((a0 + a1) + (a2 + a3));
          if(i == k) scr[0] = v;
        }
        sync();
        double vk = scr[0];
        if(vk <= 0.0)
        {
          if(fc != nullptr && tid == 0) fc[(long long)n * n] = -1.0;
          return TS_NON_POS_HESSIAN;
        }
        double lkk = sqrt(vk);
        if(i == k)
        {
          Jb[i * ldj + k] = lkk;
          ldiag[k] = lkk;
          rs[k] = 1.0 / lkk;
        }
        else if(i > k && i < n)
          Jb[i * ldj + k] = v / lkk;
        sync();
      }
      if(LPR)
      {
        const int k2 = max(k1, n - T / 4);
        bool pd = true;
#pragma unroll 1
        for(int k = k1; k < k2 && pd; ++k) pd = chol_step_split<2, true>(k);
#pragma unroll 1
        for(int k = k2; k < n && pd; ++k) pd = chol_step_split<4, true>(k);
        if(!pd)
        {
          if(fc != nullptr && tid == 0) fc[(long long)n * n] = -1.0;
          return TS_NON_POS_HESSIAN;
        }
      }
    }
    if(P.L != nullptr)
    {
      double * Lout = P.L + b * (long long)n * n;
      for(int j = 0; j < n; ++j)
        if(i < n && i >= j) Lout[i + (long long)j * n] = Jb[i * ldj + j];
    }

    // ---- J = L^-T in the upper triangle (diagonal of L kept in ldiag); L stays below for B = L^-1 N
    {
      const int j = i, jc = ic;
      const int jmax = min(n - 1, 32 * warp + 31);
      if(j < n) Jb[j * ldj + j] = rs[j];
#pragma unroll 1
      for(int r = jmax - 1; r >= 0; --r)
      {
        double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll 1
        for(int k0 = r + 1; k0 <= jmax; k0 += 4)
        {
          const double * Jk = Jb + k0 * ldj;
          const double t0 = fma(Jk[r], Jk[jc], a0);
          const double t1 = fma(Jk[ldj + r], Jk[ldj + jc], a1);
          const double t2 = fma(Jk[2 * ldj + r], Jk[2 * ldj + jc], a2);
          const double t3 = fma(Jk[3 * ldj + r], Jk[3 * ldj + jc], a3);
          a0 = k0 <= j ? t0 : a0;
          a1 = k0 + 1 <= j ? t1 : a1;
          a2 = k0 + 2 <= j ? t2 : a2;
          a3 = k0 + 3 <= j ? t3 : a3;
        }
        if(r < j && j < n) Jb[r * ldj + j] = (-((a0 + a1) + (a2 + a3))) * rs[r];
      }
    }
    sync();
    if(fc != nullptr)
    {
      // step 0 of a sequence: keep the factor for the steps that follow
#pragma unroll 4
      for(int j = 0; j < n; ++j)
        if(i < n) fc[i + (long long)j * n] = Jb[i * ldj + j];
      if(i < n)
      {
        fc[(long long)n * n + i] = ldiag[i];
        fc[(long long)n * n + n + i] = rs[i];
      }
    }
    } // (!fload)

    // ---- active normals and b_act (initializeComputationData, :383-418), then B = L^-1 N,
    //      thread = active column k
    if(i < q)
    {
      const int k = i;
      const int ci = alist[k];
      const int sv = stat[ci];
      double bk;
      if(ci < mc)
      {
        const double * cg = P.C + b * P.sC + (long long)ci * P.ldc;
        const bool neg = sv == ST_UPPER;
        for(int r = 0; r < n; ++r)
        {
          const double v = cg[r];
          *Bat(r, k) = neg ? -v : v;
        }
        bk = neg ? -bu[ci] : bl[ci];
      }
      else
      {
        const int pb = ci - mc;
        const bool neg = sv == ST_UPPER_BOUND;
        for(int r = 0; r < n; ++r) *Bat(r, k) = r == pb ? (neg ? -1.0 : 1.0) : 0.0;
        bk = neg ? -xu[pb] : xl[pb];
      }
      bact[k] = bk;
      // B(r,k) = (N(r,k) - dot4_{j<r}(L(r,j), B(j,k))) / L(r,r)
      for(int r = 0; r < n; ++r)
      {
        const double * Lr = Jb + r * ldj;
        double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        int j = 0;
        for(; j + 3 < r; j += 4)
        {
          a0 = fma(Lr[j], *Bat(j, k), a0);
          a1 = fma(Lr[j + 1], *Bat(j + 1, k), a1);
          a2 = fma(Lr[j + 2], *Bat(j + 2, k), a2);
          a3 = fma(Lr[j + 3], *Bat(j + 3, k), a3);
        }
        if(j < r) a0 = fma(Lr[j], *Bat(j, k), a0);
        if(j + 1 < r) a1 = fma(Lr[j + 1], *Bat(j + 1, k), a1);
        if(j + 2 < r) a2 = fma(Lr[j + 2], *Bat(j + 2, k), a2);
        double * br = Bat(r, k);
        *br = (*br - ((a0 + a1) + (a2 + a3))) / ldiag[r];
      }
    }
    sync();
    // L is no longer needed: clear the strict lower triangle of J
    if(i < n)
    {
#pragma unroll 4
      for(int c = 0; c < i; ++c) Jb[i * ldj + c] = 0.0;
    }

    // ---- Householder QR of B in place (unblocked), then rinv[k] = 1 / R(k,k)
#pragma unroll 1
    for(int k = 0; k < q; ++k)
    {
      const int len = n - k - 1;
      double * ess = Vp + offV(k);
      const double c0 = Rp[colR(k) + k];
      double tsq = 0.0;
      for(int t = lane; t < len; t += 32)
      {
        const double e = ess[t];
        tsq = fma(e, e, tsq);
      }
      tsq = warp_sum32(tsq);
      double tau, beta;
      const bool degenerate = tsq <= 2.2250738585072014e-308;
      if(degenerate)
      {
        tau = 0.0;
        beta = c0;
      }
      else
      {
        beta = sqrt(fma(c0, c0, tsq));
        if(c0 >= 0.0) beta = -beta;
        tau = (beta - c0) / beta;
      }
      sync(); // everybody has read column k before it is rewritten
      {
        const double den = c0 - beta;
        for(int t = tid; t < len; t += T) ess[t] = degenerate ? 0.0 : ess[t] / den;
      }
      if(tid == 0)
      {
        Rp[colR(k) + k] = beta;
        hco[k] = tau;
        rinv[k] = 1.0 / beta;
      }
      sync();
      // apply H_k to the columns j > k. JRLQP_QR4: FOUR lanes per column — lane c of a group runs chain c of the canonical dot4
      // (the entries t = c, c + 4, ... of the column, ascending), the chains are folded (a0 + a1) + (a2 + a3) by two shuffles
      // (additions commute: same bits), and the lane then updates the very entries it read — T / 4 columns per round instead of
      // one column per thread with at most q - k - 1 of the T threads busy. Same operations per entry, same order: same bits.
#if JRLQP_QR4
      if(len == 0)
      {
        for(int j = k + 1 + tid; j < q; j += T)
        {
          double * top = Rp + colR(j) + k;
          *top = *top * (1.0 - tau);
        }
      }
      else if(tau != 0.0)
      {
        const int c = tid & 3;
#pragma unroll 1
        for(int j0 = k + 1; j0 < q; j0 += T / 4) // (uniform trip count: the shuffles below are executed by every lane)
        {
          const int j = j0 + (tid >> 2);
          const bool on = j < q;
          const int jc = on ? j : q - 1;
          // rows k + 1 + t <= jc of column jc live in the packed upper storage, the rows below in the packed lower one
          const double * up = Rp + colR(jc) + k + 1;
          const double * lo = Vp + offV(jc) - (jc - k);
          const int nup = jc - k; // entries t < nup are in the upper storage
          double a = 0.0;
          for(int t = c; t < len; t += 4) a = fma(ess[t], t < nup ? up[t] : lo[t], a);
          double * top = Rp + colR(jc) + k;
          const double tp = *top;
          a = a + __shfl_xor_sync(JRLQP_FULL, a, 1);
          a = a + __shfl_xor_sync(JRLQP_FULL, a, 2);
          const double tmp = a + tp;
          __syncwarp(); // every lane of the group has read the top entry
          if(on)
          {
            if(c == 0) *top = fma(-tau, tmp, tp);
            double * upw = Rp + colR(jc) + k + 1;
            double * low = Vp + offV(jc) - (jc - k);
            for(int t2 = c; t2 < len; t2 += 4)
            {
              double * e = t2 < nup ? upw + t2 : low + t2;
              *e = fma(-(tau * ess[t2]), tmp, *e);
            }
          }
        }
      }
      sync();
#else
      for(int j = k + 1 + tid; j < q; j += T)
      {
        double * top = Rp + colR(j) + k;
        if(len == 0)
          *top = *top * (1.0 - tau);
        else if(tau != 0.0)
        {
          double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
          int t = 0;
          for(; t + 3 < len; t += 4)
          {
            a0 = fma(ess[t], *Bat(k + 1 + t, j), a0);
            a1 = fma(ess[t + 1], *Bat(k + 2 + t, j), a1);
            a2 = fma(ess[t + 2], *Bat(k + 3 + t, j), a2);
            a3 = fma(ess[t + 3], *Bat(k + 4 + t, j), a3);
          }
          if(t < len) a0 = fma(ess[t], *Bat(k + 1 + t, j), a0);
          if(t + 1 < len) a1 = fma(ess[t + 1], *Bat(k + 2 + t, j), a1);
          if(t + 2 < len) a2 = fma(ess[t + 2], *Bat(k + 3 + t, j), a2);
          const double tmp = ((a0 + a1) + (a2 + a3)) + *top;
          *top = fma(-tau, tmp, *top);
          for(int t2 = 0; t2 < len; ++t2)
          {
            double * e = Bat(k + 1 + t2, j);
            *e = fma(-(tau * ess[t2]), tmp, *e);
          }
        }
      }
      sync();
#endif
    }

    // ---- J = J Q, thread = row (rows are independent: no barrier between reflectors)
    if(i < n)
    {
      double * Jr = Jb + i * ldj;
#pragma unroll 1
      for(int k = 0; k < q; ++k)
      {
        const int len = n - k - 1;
        const double * ess = Vp + offV(k);
        const double tau = hco[k];
        if(len == 0)
          Jr[k] = Jr[k] * (1.0 - tau);
        else if(tau != 0.0)
        {
          double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
          const double * Jt = Jr + k + 1;
          int t = 0;
          for(; t + 3 < len; t += 4)
          {
            a0 = fma(Jt[t], ess[t], a0);
            a1 = fma(Jt[t + 1], ess[t + 1], a1);
            a2 = fma(Jt[t + 2], ess[t + 2], a2);
            a3 = fma(Jt[t + 3], ess[t + 3], a3);
          }
          if(t < len) a0 = fma(Jt[t], ess[t], a0);
          if(t + 1 < len) a1 = fma(Jt[t + 1], ess[t + 1], a1);
          if(t + 2 < len) a2 = fma(Jt[t + 2], ess[t + 2], a2);
          const double tmp = ((a0 + a1) + (a2 + a3)) + Jr[k];
          Jr[k] = fma(-tau, tmp, Jr[k]);
          const double tt = tau * tmp;
          for(int t2 = 0; t2 < len; ++t2) Jr[k + 1 + t2] = fma(-tt, ess[t2], Jr[k + 1 + t2]);
        }
      }
    }
    sync();

    warm_primal_dual(ab);

    // ---- constraints activated with a negative multiplier are dropped, most negative first (:83-108)
#pragma unroll 1
    for(;;)
    {
      double bu_ = -1e-14;
      int bl_ = JRLQP_NONE;
#pragma unroll
      for(int s = 0; s < W; ++s)
      {
        const int l = lane + 32 * s;
        if(l < q)
        {
          const int sv = stat[alist[l]];
          const double ul = us[l];
          if(ul < bu_ && sv != ST_FIXED && sv != ST_EQUALITY)
          {
            bu_ = ul;
            bl_ = l;
          }
        }
      }
#pragma unroll
      for(int off = 16; off >= 1; off >>= 1)
      {
        const double ou = __shfl_xor_sync(JRLQP_FULL, bu_, off);
        const int ol = __shfl_xor_sync(JRLQP_FULL, bl_, off);
        if(ou < bu_ || (ou == bu_ && ol < bl_))
        {
          bu_ = ou;
          bl_ = ol;
        }
      }
      const int lmin = __shfl_sync(JRLQP_FULL, bl_, 0);
      if(lmin == JRLQP_NONE) break;
      ++it;
      sync();
      // b_act.segment(lmin, q-1-lmin) = b_act.tail(q-1-lmin)
      if(warp == 0)
      {
        double bt[W];
#pragma unroll
        for(int s = 0; s < W; ++s)
        {
          const int k = lane + 32 * s;
          bt[s] = (k >= lmin && k + 1 < q) ? bact[k + 1] : 0.0;
        }
        __syncwarp();
#pragma unroll
        for(int s = 0; s < W; ++s)
        {
          const int k = lane + 32 * s;
          if(k >= lmin && k + 1 < q) bact[k] = bt[s];
        }
      }
      remove_constraint(lmin);
      warm_primal_dual(ab);
    }
    return TS_SUCCESS;
  }

  // ------------------------------------------------------------------------------------------
  // selectViolatedConstraint_ (src/GoldfarbIdnaniSolver.cpp:84-134).
  //
  // The scan is run by NSW "scan warps" (scan_partial), each leaving its first-minimum candidate in shared memory;
  // every thread then combines the NSW candidates after the next block barrier (scan_combine). Splitting it this way
  // lets the scan of the NEXT iteration run on warps that are otherwise parked while warp 0 does the back
  // substitution (solve()): it reads a speculative x (x + t2 z, the point of a full step) and is simply discarded
  // when the step turns out to be partial.
  //
  // Row scan (C in place / staged): the inactive constraints are first compacted into `ilist` (ballot + popc, every
  // scan warp builds the same list), then TWO lanes share one constraint: lane h = 0 owns the dot4 accumulators 0 and
  // 1 (elements k = 4j, 4j+1), lane h = 1 the accumulators 2 and 3 (k = 4j+2, 4j+3). Each accumulator is still the
  // canonical chain (k ascending), the result (a0 + a1) + (a2 + a3) is formed by one shuffle, so the bits are those of
  // dot4. A warp-level load then touches 16 rows x 32 contiguous bytes (whole sectors, nothing fetched twice), all the
  // loads of a row are in flight together (one L2 round trip per round instead of one per chunk), and a round covers
  // 16 NSW inactive constraints — the active ones (60 % at the optimum of config A) cost nothing.
  // ------------------------------------------------------------------------------------------
  static constexpr int NJB = 8; // chunks of 4 elements in flight per lane (one batch = 32 elements of a row)

  template<bool VEC>
  __device__ __forceinline__ double half_dot(const double * ci, const double * xh, const int rem, const int nj, const bool act) const
  {
    // ci, xh already point at element 2h; this lane owns the elements k' = 4j, 4j+1 < rem of that view.
    // The padding of the x vector (up to npad) is kept at zero, and a missing element is loaded as zero:
    // fma(0, 0, acc) == acc exactly (acc is never -0), so the unconditional FMAs below do not change any bit.
    double a0 = 0.0, a1 = 0.0;
#pragma unroll 1
    for(int j0 = 0; j0 < nj; j0 += NJB)
    {
      double2 v[NJB];
#pragma unroll
      for(int u = 0; u < NJB; ++u)
      {
        const int j = j0 + u;
        v[u] = make_double2(0.0, 0.0);
        if(j < nj && act)
        {
          if(VEC)
          {
            if(4 * j + 1 < rem)
              v[u] = *reinterpret_cast<const double2 *>(ci + 4 * j);
            else if(4 * j < rem)
              v[u].x = ci[4 * j];
          }
          else
          {
            if(4 * j < rem) v[u].x = ci[4 * j];
            if(4 * j + 1 < rem) v[u].y = ci[4 * j + 1];
          }
        }
      }
#pragma unroll
      for(int u = 0; u < NJB; ++u)
      {
        const int j = j0 + u;
        if(j < nj)
        {
          const double2 x2 = *reinterpret_cast<const double2 *>(xh + 4 * j);
          a0 = fma(v[u].x, x2.x, a0);
          a1 = fma(v[u].y, x2.y, a1);
        }
      }
    }
    return a0 + a1;
  }

  // One scan warp's share. xv: the point to test (16-byte aligned, zero padding); excl: a constraint to be treated as
  // active whatever its status says (the one being added by the step in flight; -1: none); sw: index of this warp
  // among the NSW scan warps.
  template<int NSW>
  __device__ __forceinline__ void scan_partial(const double * xv, const int excl, const int sw)
  {
    constexpr int TT = 32 * NSW;
    const int tt = 32 * sw + lane;
    double best = 0.0;
    double bestcx = 0.0;
    int code = JRLQP_NONE;
    bool bothneg = false;
    if(CT_SCAN && Ct != nullptr)
    {
      // wide kernels: thread = constraint over the transposed copy of C (coalesced along the constraints)
      for(int base = 0; base < mc; base += TT)
      {
        const int c = base + tt;
        const bool act = c < mc && stat[c] == ST_INACTIVE && c != excl;
        if(__ballot_sync(JRLQP_FULL, act) == 0u) continue; // warp-uniform
        const double blc = act ? bl[c] : 0.0, buc = act ? bu[c] : 0.0; // issued ahead of the dot product
        const double cx = dot4_col<CH>(Ct + ct_offset(min(c, mc - 1), n), xv, n, act);
        if(act)
        {
          const double sl = cx - blc;
          const double su = buc - cx;
          if(sl < 0.0 && su < 0.0) bothneg = true;
          if(sl < best)
          {
            best = sl;
            bestcx = cx;
            code = c * 8 + ST_LOWER;
          }
          else if(su < best)
          {
            best = su;
            bestcx = cx;
            code = c * 8 + ST_UPPER;
          }
        }
      }
    }
    else if(mc > 0)
    {
      // (1) compact list of the inactive constraints
      int ni = 0;
      for(int base = 0; base < mc; base += 32)
      {
        const int c = base + lane;
        const bool in = c < mc && stat[c] == ST_INACTIVE && c != excl;
        const unsigned msk = __ballot_sync(JRLQP_FULL, in);
        if(in) ilist[ni + __popc(msk & ((1u << lane) - 1u))] = (unsigned short)c;
        ni += __popc(msk);
      }
      __syncwarp();
      // (2) rounds of TT / 2 constraints, two lanes per constraint
      const int h = lane & 1;
      const int nj = (n + 3) >> 2;
      const int rem = n - 2 * h;
      const double * xh = xv + 2 * h;
#pragma unroll 1
      for(int base = 0; base < ni; base += TT / 2)
      {
        const int idx = base + (tt >> 1);
        const bool act = idx < ni;
        const int c = act ? (int)ilist[idx] : 0;
        const double * ci = Cb + (long long)c * ldC + 2 * h;
        const double blc = act ? bl[c] : 0.0, buc = act ? bu[c] : 0.0; // issued ahead of the dot product
        double s01;
        if(!STAGE_C && cvec)
          s01 = half_dot<true>(ci, xh, rem, nj, act);
        else
          s01 = half_dot<false>(ci, xh, rem, nj, act);
        const double s23 = __shfl_xor_sync(JRLQP_FULL, s01, 1);
        const double cx = h == 0 ? s01 + s23 : s23 + s01; // (a0 + a1) + (a2 + a3) on both lanes
        if(act)
        {
          const double sl = cx - blc;
          const double su = buc - cx;
          if(sl < 0.0 && su < 0.0) bothneg = true;
          if(sl < best)
          {
            best = sl;
            bestcx = cx;
            code = c * 8 + ST_LOWER;
          }
          else if(su < best)
          {
            best = su;
            bestcx = cx;
            code = c * 8 + ST_UPPER;
          }
        }
      }
    }
    for(int base = 0; base < nb; base += TT)
    {
      const int c = base + tt;
      if(c < nb && stat[mc + c] == ST_INACTIVE && mc + c != excl)
      {
        const double xi = xv[c];
        const double sl = xi - xl[c];
        const double su = xu[c] - xi;
        if(sl < 0.0 && su < 0.0) bothneg = true;
        if(sl < best)
        {
          best = sl;
          bestcx = xi;
          code = (mc + c) * 8 + ST_LOWER_BOUND;
        }
        else if(su < best)
        {
          best = su;
          bestcx = xi;
          code = (mc + c) * 8 + ST_UPPER_BOUND;
        }
      }
    }
    // first-minimum reduction: smallest value, ties to the smallest constraint index
#pragma unroll
    for(int off = 16; off >= 1; off >>= 1)
    {
      const double ov = __shfl_xor_sync(JRLQP_FULL, best, off);
      const double ocx = __shfl_xor_sync(JRLQP_FULL, bestcx, off);
      const int oc = __shfl_xor_sync(JRLQP_FULL, code, off);
      if(ov < best || (ov == best && oc < code))
      {
        best = ov;
        bestcx = ocx;
        code = oc;
      }
    }
    const unsigned anyneg = __ballot_sync(JRLQP_FULL, bothneg);
    // lane 0's view is published (with NaN inputs the lanes may disagree: one lane decides)
    if(lane == 0)
    {
      scr[2 + 2 * sw] = best;
      scr[3 + 2 * sw] = bestcx;
      iscr[2 * sw] = code;
      iscr[2 * sw + 1] = anyneg != 0u;
    }
  }

  // Every thread, after the barrier that follows scan_partial: the selected constraint, and its c.x for
  // computeStepLength_ (x is unchanged between the two; same dot4 => same bits). xcur: the current x, for the exact
  // sequential restatement used when some constraint has both slacks negative (bl > bu).
  template<int NSW>
  __device__ __forceinline__ Sel scan_combine(double & cx_sel)
  {
    double best = scr[2];
    double bestcx = scr[3];
    int code = iscr[0];
    int neg = iscr[1];
#pragma unroll
    for(int w = 1; w < NSW; ++w)
    {
      const double ov = scr[2 + 2 * w];
      const int oc = iscr[2 * w];
      neg |= iscr[2 * w + 1];
      if(ov < best || (ov == best && oc < code))
      {
        best = ov;
        bestcx = scr[3 + 2 * w];
        code = oc;
      }
    }
    if(neg)
    {
      Sel s = select_sequential(n, mc, nb, Cb, ldC, xs, bl, bu, xl, xu, stat);
      cx_sel = s.p < 0 ? 0.0 : (s.p < mc ? dot4_uniform(n, Cb + (long long)s.p * ldC, xs) : xs[s.p - mc]);
      return s;
    }
    cx_sel = bestcx;
    if(code == JRLQP_NONE) return {-1, ST_INACTIVE};
    return {code >> 3, code & 7};
  }

  // ------------------------------------------------------------------------------------------
  // computeStep_ (src/GoldfarbIdnaniSolver.cpp:136-148): d = J^T n+, z = J2 d2, r = R^-1 d1, and,
  // speculatively, the Givens recurrence of the addConstraint_ that may follow
  // (src/GoldfarbIdnaniSolver.cpp:226-232). Leaves d in ds, z in zs, r in rs, the rotation table in
  // gc/gs and the final rho (the new diagonal entry of R) in scr[10].
  // ------------------------------------------------------------------------------------------
  __device__ void compute_step(Sel sc)
  {
    const int j = tid;
    const int jc = min(j, n - 1);
    const bool general = sc.st < ST_LOWER_BOUND;
    // stage the selected normal once (coalesced) — it is read by d, c.z and c.x
    if(general && !cv_ready && j < n) cv[j] = Cb[(long long)sc.p * ldC + j];
    cv_ready = false;
    sync();
    PH_MARK(2); // fetch of the selected normal

    // d, thread = column
    if(general)
    {
      double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
      const double * Jc = Jb + jc;
      int i = 0;
#pragma unroll UNR_DZ
      for(; i + 3 < n; i += 4)
      {
#if JRLQP_OPT_V2
        const double2 c01 = *reinterpret_cast<const double2 *>(cv + i);
        const double2 c23 = *reinterpret_cast<const double2 *>(cv + i + 2);
        a0 = fma(Jc[i * ldj], c01.x, a0);
        a1 = fma(Jc[(i + 1) * ldj], c01.y, a1);
        a2 = fma(Jc[(i + 2) * ldj], c23.x, a2);
        a3 = fma(Jc[(i + 3) * ldj], c23.y, a3);
#else
        a0 = fma(Jc[i * ldj], cv[i], a0);
        a1 = fma(Jc[(i + 1) * ldj], cv[i + 1], a1);
        a2 = fma(Jc[(i + 2) * ldj], cv[i + 2], a2);
        a3 = fma(Jc[(i + 3) * ldj], cv[i + 3], a3);
#endif
      }
      if(i < n) a0 = fma(Jc[i * ldj], cv[i], a0);
      if(i + 1 < n) a1 = fma(Jc[(i + 1) * ldj], cv[i + 1], a1);
      if(i + 2 < n) a2 = fma(Jc[(i + 2) * ldj], cv[i + 2], a2);
      double dj = (a0 + a1) + (a2 + a3);
      if(sc.st == ST_UPPER) dj = -dj;
      if(j < n) ds[j] = dj;
    }
    else
    {
      const double * Jrow = Jb + (sc.p - mc) * ldj;
      if(j < n) ds[j] = sc.st == ST_UPPER_BOUND ? -Jrow[j] : Jrow[j];
    }
    sync();
    PH_MARK(3); // d

    // (seeds of the Givens recurrence, part 1: suffix sums of d_j^2 inside the warp, warp totals to shared memory)
    double sfx = 0.0, dj_seed = 0.0;
    if(SEEDS_IN_Z)
    {
      dj_seed = (j >= q && j < n) ? ds[j] : 0.0;
      sfx = dj_seed * dj_seed;
#pragma unroll
      for(int off = 1; off < 32; off <<= 1)
      {
        const double t = __shfl_down_sync(JRLQP_FULL, sfx, off);
        if(lane + off < 32) sfx = sfx + t;
      }
      if(lane == 0) scr[2 + warp] = sfx;
    }
    // z, thread = row: z[i] = dot4_{j=q..n-1}(J(i,j), d[j]), accumulator (j-q)&3
    {
      double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
      const double * Jr = Jb + jc * ldj;
      int c = q;
#pragma unroll UNR_DZ
      for(; c + 3 < n; c += 4)
      {
        a0 = fma(Jr[c], ds[c], a0);
        a1 = fma(Jr[c + 1], ds[c + 1], a1);
        a2 = fma(Jr[c + 2], ds[c + 2], a2);
        a3 = fma(Jr[c + 3], ds[c + 3], a3);
      }
      if(c < n) a0 = fma(Jr[c], ds[c], a0);
      if(c + 1 < n) a1 = fma(Jr[c + 1], ds[c + 1], a1);
      if(c + 2 < n) a2 = fma(Jr[c + 2], ds[c + 2], a2);
      if(j < n) zs[j] = (a0 + a1) + (a2 + a3);
    }

    // the two serial recurrences, concurrently on different warps when W > 1
    sync();
    PH_MARK(4); // z
    if(SEEDS_IN_Z)
    {
      // part 2: S_j = d_j^2 + d_{j+1}^2 + ... ~ rho_j^2, then the seeds of link j - 1 (see givens_recurrence); the chain warp
      // alone waits for them (named barrier 3: the other warps only arrive)
      double S = sfx;
      for(int w = warp + 1; w < W; ++w) S = S + scr[2 + w];
      double rsd = rsqrt_seed(S);
      if(j == n - 1) rsd = dj_seed < 0.0 ? -rsd : rsd;
      const double t = ds[j - 1] * rsd; // (j == 0: an unused read of the padding before d)
      const double ss = fma(t, t, 1.0);
      const double ysd = rsqrt_seed(ss);
      if(W >= JRLQP_CHAIN_BF_MINW && JRLQP_CHAIN_BF)
      {
        // link j - 1 is PREDICTED off the fast branch of makeGivens (|p| <= |rho|, p != 0) when the seeds say |p / rho| is
        // not safely below 1; the recurrence runs branch-free below the lowest such link (and checks the exact test there)
        const bool ps = j > q && j < n && (!(fabs(t) <= 0.99999) || ds[j - 1] == 0.0);
        const unsigned mk = __ballot_sync(JRLQP_FULL, ps);
        if(lane == 0) reinterpret_cast<int *>(scr + 6)[warp] = mk ? 32 * warp + __ffs(mk) - 2 : JRLQP_NONE;
      }
      if(j >= q && j < n)
      {
        if(j > q)
        {
          grec[2 * (j - 1)] = make_double2(rsd, ss * ysd);
          double2 hk = make_double2(0.5 * ysd, 0.0);
          rec_kind(hk) = 3 | 8;
          grec[2 * (j - 1) + 1] = hk;
        }
        else
          scr[11] = rsd;
      }
      if(warp == (W > 1 ? 1 : 0))
        asm volatile("bar.sync 3, %0;" ::"n"(32 * W) : "memory");
      else
        asm volatile("bar.arrive 3, %0;" ::"n"(32 * W) : "memory");
    }
  }

  // r = R^-1 d(0:q): column-oriented back substitution with true division (r_k = w_k / R(k,k), then
  // w_i = fma(-r_k, R(i,k), w_i) for i < k, k descending), one warp, rows over the lanes (W slots).
  //
  // What bounds it is the latency of one link, so the loop carries only what is serial:
  //  * the pivot value is kept UNIFORM: while link k is being divided, every lane already holds w_{k-1} as it was
  //    before link k (one shuffle, issued ahead of the quotient), and applies link k's update to it itself
  //    (wp = fma(-r_k, R(k-1,k), w_{k-1}) — the very operation lane k-1 performs on its own row, hence the same
  //    bits). The dependent chain of a link is the quotient (3 FMAs) + 1 FMA; the shuffle is off the chain;
  //  * the quotient comes from the reciprocal kept in rinv[k] (div_rcp), and its proof of correct rounding is NOT
  //    evaluated in the loop: lane k's row is never touched after its pivot, so after the loop lane k still holds
  //    the dividend w_k, redoes its own quotient (same inputs, same bits) and checks the proof — all links at
  //    once. If one proof is declined (rare) the pass is redone with the stock division.
  __device__ __forceinline__ void back_substitution()
  {
    double w[W], rr[W];
#pragma unroll
    for(int s = 0; s < W; ++s) w[s] = lane + 32 * s < q ? ds[lane + 32 * s] : 0.0;
    if(q > 0)
    {
      int k = q - 1;
      const double * Rk = Rp + colR(k);
      double wp = ds[k];
      if(W >= JRLQP_BS_PREFETCH)
      {
        // wide kernels (one CTA per SM, pure latency): the operands of a link are loaded one link ahead (their
        // addresses do not depend on the recurrence); two links per trip so that no register is copied
        double rkk = Rk[k], ri = rinv[k];
        double rnk = Rk[k - 1]; // R(k-1, k); k == 0: an unused read of the word before R
        double col[W];
#pragma unroll
        for(int s = 0; s < W; ++s) col[s] = Rk[lane + 32 * s]; // rows >= k: an unused read past the column
#pragma unroll 2
        for(; k >= 0; --k)
        {
          const double * Rn = Rk - k; // column k - 1 (k == 0: unused reads around the first column)
          const double rkk_n = Rn[k - 1], ri_n = rinv[k - 1], rnk_n = Rn[k - 2];
          double col_n[W];
#pragma unroll
          for(int s = 0; s < W; ++s) col_n[s] = Rn[lane + 32 * s];
          const double wn = __shfl_sync(JRLQP_FULL, pick<W>(w, (k - 1) >> 5), (k - 1) & 31);
          const double q0 = wp * ri;
          const double e = fma(-rkk, q0, wp);
          const double rk = fma(e, ri, q0);
#pragma unroll
          for(int s = 0; s < W; ++s)
            if(lane + 32 * s < k) w[s] = fma(-rk, col[s], w[s]);
          wp = fma(-rk, rnk, wn);
          Rk = Rn;
          rkk = rkk_n;
          ri = ri_n;
          rnk = rnk_n;
#pragma unroll
          for(int s = 0; s < W; ++s) col[s] = col_n[s];
        }
      }
      else
      {
        // narrow kernels (several CTAs per SM, instruction-cache sensitive): the smallest loop body
#pragma unroll 1
        for(; k >= 0; --k)
        {
          const double rkk = Rk[k], ri = rinv[k];
          const double rnk = Rk[k - 1]; // R(k-1, k); k == 0: an unused read of the word before R
          double col[W];
#pragma unroll
          for(int s = 0; s < W; ++s) col[s] = Rk[lane + 32 * s]; // rows >= k: an unused read past the column
          const double wn = __shfl_sync(JRLQP_FULL, pick<W>(w, (k - 1) >> 5), (k - 1) & 31);
          const double q0 = wp * ri;
          const double e = fma(-rkk, q0, wp);
          const double rk = fma(e, ri, q0);
#pragma unroll
          for(int s = 0; s < W; ++s)
            if(lane + 32 * s < k) w[s] = fma(-rk, col[s], w[s]);
          wp = fma(-rk, rnk, wn);
          Rk -= k;
        }
      }
    }
    // every lane redoes the quotient of its own row (same inputs, same bits) and checks the proof
    bool ok = true;
#pragma unroll
    for(int s = 0; s < W; ++s)
    {
      const int r = lane + 32 * s;
      rr[s] = 0.0;
      if(r < q)
      {
        bool okk;
        rr[s] = div_rcp(w[s], Rp[colR(r) + r], rinv[r], okk);
        ok = ok && okk;
      }
    }
    if(__all_sync(JRLQP_FULL, ok))
    {
#pragma unroll
      for(int s = 0; s < W; ++s)
        if(lane + 32 * s < q) rs[lane + 32 * s] = rr[s];
    }
    else
      back_substitution_exact<W>(q, lane, ds, Rp, rs); // a proof was declined: every quotient by the stock division
  }

  // Givens recurrence of the add that may follow: the algorithm of givens_chain above (seeds, 7-operation link,
  // deferred proofs), with the per-link data in ONE array of 32-byte records so that the loop runs on two pointers
  // with immediate offsets: grec[2i] = (rs_i, then t_i, then c_i ; ys_i, then u_i, then s_i),
  // grec[2i+1] = (h_i, then the incoming rho ; branch taken | 8 if the link is to be proven), where the seeds are
  // rs ~ 1 / rho, g ~ u, h ~ 1 / (2 u). Every lane stores (same value, same address): no predicate in the loop.
  static constexpr bool SPLIT_CHAIN = JRLQP_OPT_CS && W > 1;
  static constexpr bool SEEDS_IN_Z = JRLQP_SEEDS_IN_Z && W > 1;
  static constexpr int CHAIN_UNR = W >= 3 ? 2 : JRLQP_CHAIN_UNR;
  static constexpr bool CHAIN_VOTE = W >= 3 ? false : (JRLQP_CHAIN_VOTE != 0); // measured: profiles/r02j_ab_*.txt
  __device__ __forceinline__ static int & rec_kind(double2 & r) { return reinterpret_cast<int *>(&r.y)[0]; }
  __device__ __forceinline__ void givens_recurrence()
  {
    double2 * const rec = grec;
    // ---- seeds: element j = top - lane of d, chunks of 32 from the last element down; suffix sums by a warp scan
    //      (SEEDS_IN_Z: already computed by all the threads in compute_step)
    if(!SEEDS_IN_Z)
    {
      double carry = 0.0;
#pragma unroll 1
      for(int top = n - 1; top >= q; top -= 32)
      {
        const int j = top - lane;
        const bool valid = j >= q;
        const double dj = valid ? ds[j] : 0.0;
        const double dm = valid ? ds[j - 1] : 0.0; // d of the link this element feeds (j == 0: an unused read of the padding before d)
        double sc = dj * dj;
#pragma unroll
        for(int off = 1; off < 32; off <<= 1)
        {
          const double t = __shfl_up_sync(JRLQP_FULL, sc, off);
          if(lane >= off) sc = sc + t;
        }
        const double S = carry + sc;
        carry = __shfl_sync(JRLQP_FULL, S, 31);
        double rsd = rsqrt_seed(S);
        if(j == n - 1) rsd = dj < 0.0 ? -rsd : rsd; // the first rho is d[n-1] itself, sign included
        const double t = dm * rsd;
        const double ss = fma(t, t, 1.0);
        const double ysd = rsqrt_seed(ss);
        if(valid)
        {
          if(j > q)
          {
            rec[2 * (j - 1)] = make_double2(rsd, ss * ysd); // ~ 1 / rho, ~ u = sqrt(1 + t^2)
            double2 hk = make_double2(0.5 * ysd, 0.0); // ~ 1 / (2 u)
            rec_kind(hk) = 3 | 8;
            rec[2 * (j - 1) + 1] = hk;
          }
          else
            scr[11] = rsd; // ~ 1 / rho_q: reciprocal of the new diagonal entry of R (an approximation is all rinv needs)
        }
      }
      __syncwarp();
    }
    bool careful = false;
    constexpr bool WIDE_BF = W >= JRLQP_CHAIN_BF_MINW && SEEDS_IN_Z && JRLQP_CHAIN_BF != 0;
    constexpr int BF_UNR = JRLQP_CHAIN_BF_UNR;
    int jclean = JRLQP_NONE; // every link below this one is predicted on the fast branch of makeGivens
    if(WIDE_BF)
    {
      const int * jc = reinterpret_cast<const int *>(scr + 6);
      jclean = jc[0];
#pragma unroll
      for(int w = 1; w < W; ++w) jclean = min(jclean, jc[w]);
    }
#pragma unroll 1
    for(;;)
    {
      double rho = ds[n - 1];
      int i = n - 2;
      const double * pd = ds + i;
      double2 * pr = rec + 2 * i;
      double p = pd[0]; // (n == 1: unused reads of the padding stored before the vectors)
      double2 sd = pr[0]; // (~ 1 / rho, ~ u)
      double hh = pr[1].x; // ~ 1 / (2 u)
      double q0s = p * sd.x; // first product of the quotient p / rho, issued ahead
      const int ia = (WIDE_BF && !careful) ? max(q, jclean) : q; // the links i >= ia keep the branch
#pragma unroll CHAIN_UNR
      for(; i >= ia; --i)
      {
        // operands of the next link (i == 0: unused reads of the padding)
        const double pn = pd[-1];
        const double2 sn = pr[-2];
        const double hn = pr[-1].x;
        // ---- fast path, 6 dependent operations: t = p / rho (one correction of p rs), s = 1 + t^2,
        //      u = sqrt(s) (one correction of the seed: u = g + (s - g^2) / (2 g)), r = |rho| u
        const double e3 = fma(-rho, q0s, p);
        double a = fma(e3, sd.x, q0s);
        const double s = fma(a, a, 1.0);
        const double rem = fma(-sd.y, sd.y, s);
        const double us = fma(rem, hh, sd.y);
        double r = fabs(rho) * us; // == rho * u with u = sign(rho) us, bit for bit
        double u = __hiloint2double(__double2hiint(us) | (__double2hiint(rho) & 0x80000000), __double2loint(us));
        const bool slow = careful || !(fabs(p) <= fabs(rho)) || p == 0.0;
        if(CHAIN_VOTE ? __any_sync(JRLQP_FULL, slow) : slow) // (uniform: every lane holds the same values)
        {
          int kd;
          if(rho == 0.0) // the trivial branches of makeGivens in line (runs of exact zeros in structured problems)
          {
            kd = 0;
            a = p;
            u = 1.0;
            r = fabs(p);
          }
          else if(p == 0.0)
          {
            kd = 1;
            a = rho;
            u = 1.0;
            r = fabs(rho);
          }
          else
          {
            const GivensLink gl = givens_link_slow(p, rho, 0.0, false);
            a = gl.a;
            u = gl.u;
            r = gl.r;
            kd = gl.kind;
          }
          rec_kind(pr[1]) = kd;
        }
        pr[0] = make_double2(a, u);
        pr[1].x = rho;
        rho = r;
        p = pn;
        sd = sn;
        hh = hn;
        q0s = pn * sn.x;
        --pd;
        pr -= 2;
      }
      if(WIDE_BF)
      {
        // ---- the links predicted on the fast branch: the same six operations, no branch; the operands are loaded two
        //      links ahead so that the first product of the next quotient is ready when its link starts
        bool bad = false;
        double pn = pd[-1];
        double2 sn = pr[-2];
        double hn = pr[-1].x;
#pragma unroll BF_UNR
        for(; i >= q; --i)
        {
          const double p2 = pd[-2]; // (reads below link q: addressable shared memory, values unused)
          const double2 s2 = pr[-4];
          const double h2 = pr[-3].x;
          const double q0n = pn * sn.x;
          const double e3 = fma(-rho, q0s, p);
          const double a = fma(e3, sd.x, q0s);
          const double s = fma(a, a, 1.0);
          const double rem = fma(-sd.y, sd.y, s);
          const double us = fma(rem, hh, sd.y);
          const double r = fabs(rho) * us;
          const double u = __hiloint2double(__double2hiint(us) | (__double2hiint(rho) & 0x80000000), __double2loint(us));
          bad = bad || !(fabs(p) <= fabs(rho)) || p == 0.0;
          pr[0] = make_double2(a, u);
          pr[1].x = rho;
          rho = r;
          p = pn;
          sd = sn;
          hh = hn;
          q0s = q0n;
          pn = p2;
          sn = s2;
          hn = h2;
          --pd;
          pr -= 2;
        }
        if(bad)
        {
          // a link the seeds had predicted on the fast branch was not (never seen; the seeds are accurate to ~ 2^-40 and the
          // prediction keeps a margin of 1e-5): the chain again, every link taken literally (the seeds are overwritten)
          careful = true;
          continue;
        }
      }
      scr[10] = rho;
      __syncwarp();
      if(careful) break;
      // ---- deferred proofs, one link per lane
      bool ok = true;
#pragma unroll 1
      for(int j = q + lane; j <= n - 2; j += 32)
      {
        const double2 tu = rec[2 * j];
        double2 rk = rec[2 * j + 1];
        if(rec_kind(rk) & 8)
        {
          const double a = tu.x, rh = rk.x, us = fabs(tu.y);
          // t == RN(p / rho) iff |p - rho t| < |rho| ulp(t) / 2, t not a power of two, nothing near underflow
          const double e2 = fma(-rh, a, ds[j]); // exact remainder
          const int ahi = __double2hiint(a);
          const double hu = __hiloint2double((ahi & 0x7ff00000) - 0x03500000, 0);
          const double tol = fabs(rh) * hu;
          const bool pow2 = ((ahi & 0xfffff) | __double2loint(a)) == 0;
          // u == RN(sqrt(s)), s = 1 + t^2 in [1, 2] (so ulp(u) = 2^-52): |s - u^2| < u 2^-52 (1 - 2^-41)
          const double s = fma(a, a, 1.0);
          const double rem2 = fma(-us, us, s);
#ifndef JRLQP_DIAG_NOPROOF
          ok = ok && fabs(e2) < tol && tol > 1e-270 && !pow2 && fabs(rem2) < us * 0x1.ffffffffffp-53 && us >= 1.0 && us < 2.0;
#endif
        }
      }
      if(__all_sync(JRLQP_FULL, ok)) break;
      careful = true;
    }
    if(!SPLIT_CHAIN)
    {
#pragma unroll 1
      for(int i = q + lane; i <= n - 2; i += 32) rotation_cs(i);
    }
  }
  // (c, s) of rotation i from what the recurrence stashed (second division of makeGivens), in place
  __device__ __forceinline__ void rotation_cs(const int i)
  {
    const double2 tu = grec[2 * i];
    double2 rk = grec[2 * i + 1];
    grec[2 * i] = givens_cs(rec_kind(rk), tu.x, tu.y);
  }

  // ------------------------------------------------------------------------------------------
  // computeStepLength_ (src/GoldfarbIdnaniSolver.cpp:150-219), incl. the activationStatus(k) quirk, in two halves
  // that different warps can evaluate concurrently: len_t1 needs r (after the back substitution), len_z only z and x.
  // ------------------------------------------------------------------------------------------
  // t1: first minimum of u[k]/r[k] over r[k] > 0 and status_[k] not in {EQUALITY, FIXED}; one warp.
  __device__ __forceinline__ void len_t1(double & t1, int & l, const bool need_t1)
  {
    const double big = P.big_bnd;
    t1 = big;
    l = 0;
    // (addInitialConstraint takes the exact step onto the constraint: no ratio test, src/GoldfarbIdnaniSolver.cpp:295-338)
    if(need_t1 || !(JRLQP_OPT_T1 && W == 2)) // measured: +1.0 % at W = 2, -2.2 % at W = 1, -0.7 % at W = 4 (profiles/r01z_ab_*.txt)
    {
      double bt = big;
      int bl_ = JRLQP_NONE;
#pragma unroll
      for(int s = 0; s < W; ++s)
      {
        int k = lane + 32 * s;
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
      if(bl_ != JRLQP_NONE)
      {
        t1 = bt;
        l = bl_;
      }
    }
  }

  // t2 = (b - c.x) / (c.z) if ||z|| > 1e-14, nz = ConstraintNormal::dot(z) (src/GoldfarbIdnaniSolver.cpp:289-293); one warp.
  __device__ __forceinline__ void len_z(Sel sc, bool cx_valid, double cx_in, double & t2, double & nz, bool & zpos)
  {
    // ||z|| with the dot32 order: lane accumulates k = lane, lane+32, ... then the xor butterfly
    double zz = 0.0;
    for(int k = lane; k < n; k += 32)
    {
      double zk = zs[k];
      zz = fma(zk, zk, zz);
    }
    zpos = sqrt(warp_sum32(zz)) > 1e-14;

    t2 = P.big_bnd;
    double cz, num;
    if(sc.st < ST_LOWER_BOUND)
    {
      // dot4(c, z) and dot4(c, x): accumulator t = sum over k = t, t+4, ... (ascending) is an
      // independent chain, so lane t runs chain t of c.z and lane 4+t chain t of c.x; the result is
      // (a0 + a1) + (a2 + a3), exactly the canonical dot4.
      const double * v = (lane & 4) ? xs : zs;
      double acc = 0.0;
      if(lane < 8)
      {
#pragma unroll 2
        for(int k = lane & 3; k < n; k += 4) acc = fma(cv[k], v[k], acc);
      }
      const double a1 = __shfl_down_sync(JRLQP_FULL, acc, 1);
      const double s01 = acc + a1; // valid on lanes 0, 2, 4, 6
      const double s23 = __shfl_down_sync(JRLQP_FULL, s01, 2);
      const double dot = s01 + s23; // valid on lanes 0 (c.z) and 4 (c.x)
      cz = __shfl_sync(JRLQP_FULL, dot, 0);
      const double cxn = __shfl_sync(JRLQP_FULL, dot, 4);
      nz = sc.st == ST_UPPER ? -cz : cz;
      const double b = sc.st == ST_UPPER ? bu[sc.p] : bl[sc.p]; // EQUALITY: bl (addInitialConstraint)
      num = b - (cx_valid ? cx_in : cxn);
    }
    else
    {
      const int pb = sc.p - mc;
      cz = zs[pb];
      nz = sc.st == ST_UPPER_BOUND ? -cz : cz;
      const double b = sc.st == ST_UPPER_BOUND ? xu[pb] : xl[pb];
      num = b - xs[pb];
    }
    if(zpos) t2 = num / cz;
  }

  // x += t z ; f += t (n+.z) (t/2 + u[q]) ; u(0:q) -= t r ; u[q] += t   — by warp 0 alone (rows and
  // multipliers over its lanes, W slots), right after step_length on the same warp.
  __device__ __forceinline__ void take_step(double t, double nz, bool primal)
  {
    const double uq = us[q];
    __syncwarp(); // every lane has finished reading x, u, z, r in step_length
    if(primal)
    {
#pragma unroll
      for(int s = 0; s < W; ++s)
      {
        const int r = lane + 32 * s;
        if(r < n) xs[r] = fma(t, zs[r], xs[r]);
      }
      f += (t * nz) * (0.5 * t + uq);
    }
#pragma unroll
    for(int s = 0; s < W; ++s)
    {
      const int k = lane + 32 * s;
      if(k < q) us[k] = fma(-t, rs[k], us[k]);
    }
    if(lane == 0) us[q] = uq + t;
    __syncwarp();
  }

  // ------------------------------------------------------------------------------------------
  // addConstraint (src/DualSolver.cpp:231-235) + addConstraint_ (src/GoldfarbIdnaniSolver.cpp:221-237):
  // apply the rotation table built by givens_recurrence(), thread = row of J, the rotated value
  // of column i+1 is carried in a register from one rotation to the next.
  // ------------------------------------------------------------------------------------------
  __device__ void add_constraint()
  {
    q += 1; // the active list / status entry was written by warp 0 (solve())
    if(tid < n)
    {
      double * Jr = Jb + tid * ldj;
      if(q - 1 <= n - 2)
      {
        double y = Jr[n - 1];
        int i = n - 2;
        const int lo = q - 1;
#if JRLQP_OPT_ADD
        // chunks of PF rotations, two register sets (A, B) used alternately: the operands of the next chunk are
        // loaded BEFORE the results of the current one are stored (the compiler cannot move a load above a store it
        // cannot disambiguate), and no register is copied from one set to the other
        double xa[PF], xb[PF];
        double2 ca[PF], cb[PF];
#  define JRLQP_ROT_LOAD(X, Cc, at)            \
    _Pragma("unroll") for(int u = 0; u < PF; ++u) \
    {                                          \
      X[u] = Jr[(at) - u];                     \
      Cc[u] = grec[2 * ((at) - u)];            \
    }
#  define JRLQP_ROT_APPLY(X, Cc)                               \
    {                                                          \
      double o[PF];                                            \
      _Pragma("unroll") for(int u = 0; u < PF; ++u)            \
      {                                                        \
        const double c = Cc[u].x, sn = Cc[u].y, xi = X[u];     \
        o[u] = fma(c, y, sn * xi);                             \
        y = fma(c, xi, -(sn * y));                             \
      }                                                        \
      _Pragma("unroll") for(int u = 0; u < PF; ++u) Jr[i - u + 1] = o[u]; \
      i -= PF;                                                 \
    }
        if(i - (PF - 1) >= lo)
        {
          JRLQP_ROT_LOAD(xa, ca, i)
#pragma unroll 1
          for(;;)
          {
            const bool moreB = i - (2 * PF - 1) >= lo;
            if(moreB) { JRLQP_ROT_LOAD(xb, cb, i - PF) }
            JRLQP_ROT_APPLY(xa, ca)
            if(!moreB) break;
            const bool moreA = i - (2 * PF - 1) >= lo;
            if(moreA) { JRLQP_ROT_LOAD(xa, ca, i - PF) }
            JRLQP_ROT_APPLY(xb, cb)
            if(!moreA) break;
          }
        }
#  undef JRLQP_ROT_LOAD
#  undef JRLQP_ROT_APPLY
#else
        // chunks of PF rotations; the operands of the next chunk are loaded BEFORE the results of the
        // current one are stored (the compiler cannot move a load above a store it cannot disambiguate)
        double xv[PF];
        double2 cv4[PF];
        if(i - (PF - 1) >= lo)
        {
#pragma unroll
          for(int u = 0; u < PF; ++u)
          {
            xv[u] = Jr[i - u];
            cv4[u] = grec[2 * (i - u)];
          }
        }
#pragma unroll 1
        while(i - (PF - 1) >= lo)
        {
          double xn[PF];
          double2 cn[PF];
          const bool more = i - (2 * PF - 1) >= lo;
          if(more)
          {
#pragma unroll
            for(int u = 0; u < PF; ++u)
            {
              xn[u] = Jr[i - PF - u];
              cn[u] = grec[2 * (i - PF - u)];
            }
          }
          double o[PF];
#pragma unroll
          for(int u = 0; u < PF; ++u)
          {
            const double c = cv4[u].x, sn = cv4[u].y, xi = xv[u];
            o[u] = fma(c, y, sn * xi);
            y = fma(c, xi, -(sn * y));
          }
#pragma unroll
          for(int u = 0; u < PF; ++u) Jr[i - u + 1] = o[u];
          i -= PF;
          if(more)
          {
#pragma unroll
            for(int u = 0; u < PF; ++u)
            {
              xv[u] = xn[u];
              cv4[u] = cn[u];
            }
          }
        }
#endif
#pragma unroll 1
        for(; i >= lo; --i)
        {
          const double2 cs2 = grec[2 * i];
          const double c = cs2.x, sn = cs2.y;
          const double xi = Jr[i];
          Jr[i + 1] = fma(c, y, sn * xi);
          y = fma(c, xi, -(sn * y));
        }
        Jr[q - 1] = y;
      }
      // R(0:q, q-1) = d(0:q), with d[q-1] = rho
      if(tid < q) Rp[colR(q - 1) + tid] = tid == q - 1 ? scr[10] : ds[tid];
      if(tid == q - 1) rinv[tid] = scr[11];
    }
    sync();
  }

  // ------------------------------------------------------------------------------------------
  // removeConstraint (src/DualSolver.cpp:237-244) + removeConstraint_ (src/GoldfarbIdnaniSolver.cpp:239-256).
  // Rare (about one per solve): done by warp 0 alone, rows / columns over its lanes (W slots).
  // ------------------------------------------------------------------------------------------
  __device__ void remove_constraint(int l)
  {
    sync();
    if(warp == 0)
    {
      // u.segment(l, q-l) = u.tail(q-l) (u has q+1 entries) ; A_.deactivate(l)
      double ut[W];
      int at[W];
#pragma unroll
      for(int s = 0; s < W; ++s)
      {
        int k = lane + 32 * s;
        ut[s] = (k >= l && k < q) ? us[k + 1] : 0.0;
        at[s] = (k >= l && k + 1 < q) ? alist[k + 1] : -1;
      }
      int removed = alist[l];
      __syncwarp();
#pragma unroll
      for(int s = 0; s < W; ++s)
      {
        int k = lane + 32 * s;
        if(k >= l && k < q) us[k] = ut[s];
        if(k >= l && k + 1 < q) alist[k] = at[s];
      }
      if(lane == 0) stat[removed] = ST_INACTIVE;
      const int qn = q - 1;
      __syncwarp();
#pragma unroll 1
      for(int i = l; i < qn; ++i)
      {
        double * Ri = Rp + colR(i);
        double * Ri1 = Rp + colR(i + 1);
        // R.col(i).head(i) = R.col(i+1).head(i)
#pragma unroll
        for(int s = 0; s < W; ++s)
        {
          int k = lane + 32 * s;
          if(k < i) Ri[k] = Ri1[k];
        }
        double c, sn, r;
        make_givens(Ri1[i], Ri1[i + 1], c, sn, r);
        __syncwarp();
        if(lane == 0)
        {
          Ri[i] = r;
          rinv[i] = 1.0 / r;
        }
        // rows i, i+1 of columns i+2 .. q (lane = column)
#pragma unroll
        for(int s = 0; s < W; ++s)
        {
          int j = i + 2 + lane + 32 * s;
          if(j <= qn)
          {
            double * Rj = Rp + colR(j);
            double xi = Rj[i], yi = Rj[i + 1];
            Rj[i] = fma(c, xi, -(sn * yi));
            Rj[i + 1] = fma(c, yi, sn * xi);
          }
        }
        // columns i, i+1 of J (lane = row)
#pragma unroll
        for(int s = 0; s < W; ++s)
        {
          int rrow = lane + 32 * s;
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
    q -= 1;
    sync();
  }

  // ------------------------------------------------------------------------------------------
  // DualSolver::solve (src/DualSolver.cpp:91-168) for problem b, with initActiveSet /
  // addInitialConstraint (src/GoldfarbIdnaniSolver.cpp:268-338) folded into the same loop: the
  // pre-activation of an equality (or fixed variable) is a step whose constraint is given instead
  // of selected, whose length is the exact step onto the constraint, and which always ends with an
  // add. One loop body => one copy of every phase in the instruction stream (the kernel is
  // instruction-cache sensitive, see profiles/r01b_*).
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
    cvec = ((reinterpret_cast<unsigned long long>(Cb) & 15ull) == 0ull) && ((ldC & 1) == 0);

    PH_DECL;
    stage_ct(b);
    cv_ready = false;
    int it = 0;
    int cursor = 0; // next constraint / bound to test for pre-activation; m when that phase is over
    if(WARM)
    {
      // experimental::GoldfarbIdnaniSolver: the equalities are part of the initial factorisation
      const int st0 = init_warm(b, it);
      if(st0 != TS_SUCCESS)
      {
        write_failure(b, st0);
        return;
      }
      cursor = m;
    }
    else
    {
      const bool init_ok = init(b);
      if(!init_ok)
      {
        write_failure(b, TS_NON_POS_HESSIAN);
        return;
      }
    }
    PH_MARK(0);

    int status = TS_MAX_ITER_REACHED;
    bool skip = false;
    bool have_sel = false; // the constraint of this iteration was already selected during the previous one
    Sel sc{-1, ST_INACTIVE};
    double cx_sel = 0.0;
    const double big = P.big_bnd;
    // Warp roles inside a step: warp 0 runs the back substitution, the ratio test, the decision and the
    // step; warp CW the Givens recurrence of the add that may follow. The two recurrences are about equally long, so
    //  * W <= 2: the z-dependent half of the step length stays on warp 0 and the next constraint scan is done by all
    //    the threads once the step is taken (x final);
    //  * W >= 3 (SPEC): warps [2, W) have no recurrence to run: they evaluate the z-dependent half of the step length
    //    and then — speculatively, at the point x + t2 z a full step would reach — the constraint scan of the NEXT
    //    iteration, concurrently with the recurrences; the result is simply discarded when the step turns out to be
    //    partial.
    constexpr bool SPEC = W >= 3;
    constexpr int CW = W > 1 ? 1 : 0;
    constexpr int SW = SPEC ? 2 : 0; // first scan warp
    constexpr int NSW = SPEC ? W - 2 : W; // number of scan warps
    int * dec = iscr + 2 * W; // [0] add, [1] l, [2] status on break (-1: none), [3] ||z|| > 1e-14, [4] speculative scan done, [5] the next selection is wanted
    double * decd = scr + 13; // [0] f, [1] t2, [2] n+.z
    double * xs2 = cv; // speculative x: the staged normal is dead once len_z has read it
#pragma unroll 1
    for(;;)
    {
      bool pre = false;
      // initActiveSet: equalities (bl == bu) in order, then fixed variables (xl == xu) in order
      while(cursor < m)
      {
        const int c = cursor++;
        if(eqf[c])
        {
          sc = {c, c < mc ? ST_EQUALITY : ST_FIXED};
          pre = true;
          break;
        }
      }
      bool sel_only = false; // this pass only selects (no selection was made ahead: first iteration without equalities)
      if(pre && q >= n)
      {
        // more than nbVar equalities / fixed variables: the reference would write past its workspaces; the status is
        // the one its experimental solver returns for the same input
        write_failure(b, TS_OVERCONSTRAINED_PROBLEM);
        return;
      }
      if(!pre)
      {
        if(it >= P.max_iter) break; // MAX_ITER_REACHED
        if(!skip)
        {
          if(!have_sel)
            sel_only = true;
          else if(sc.st == ST_INACTIVE)
          {
            status = TS_SUCCESS;
            break;
          }
        }
      }
      bool add = false;
      int l = 0;
      int next_pre = -1; // the next constraint initActiveSet will pre-activate (-1: none left)
      bool scan_now = sel_only; // the scan warps run the scan in this pass
      const double * xv = xs; // ... of this point
      int excl = -1;
      if(!sel_only)
      {
        // will the step after this one need a fresh selection (i.e. is the pre-activation phase over)?
        for(int c = cursor; c < m; ++c)
          if(eqf[c])
          {
            next_pre = c;
            break;
          }
        const bool more_pre = next_pre >= 0;
        const bool want_next = !more_pre && (pre || it + 1 < P.max_iter);
        if(pre || !skip)
        {
          if(tid == 0) us[q] = 0.0; // published by the barriers of compute_step
        }
        PH_MARK(11);
        compute_step(sc);
        // the warp that evaluates the z-dependent half of the step length
        const int zw = SPEC ? (want_next ? SW : 0) : ((JRLQP_ZPART_W1 && W == 2 && !pre) ? 1 : 0);
        double t2 = 0.0, nz = 0.0;
        bool zpos = false;
        if(warp == zw)
        {
          len_z(sc, !pre && !skip, cx_sel, t2, nz, zpos);
          PH_MARK(7);
          if(W > 1 && zw != 0)
          {
            const double ts = pre ? (zpos ? t2 : 0.0) : t2; // the length of a full step (pre-activation: the exact step)
            const bool spec = SPEC && ts < big;
            if(SPEC)
            {
              __syncwarp(); // cv has been read by every lane
#pragma unroll
              for(int s = 0; s < W; ++s)
              {
                const int r = lane + 32 * s;
                if(r < n) xs2[r] = fma(ts, zs[r], xs[r]);
              }
            }
            if(lane == 0)
            {
              decd[1] = t2;
              decd[2] = nz;
              dec[3] = zpos;
              dec[4] = spec;
            }
            asm volatile("bar.arrive 1, 64;" ::: "memory"); // hand t2, n+.z over to warp 0 (which waits with bar.sync 1)
            if(SPEC && NSW > 1) asm volatile("bar.sync 2, %0;" ::"n"(32 * (NSW > 1 ? NSW : 1)) : "memory"); // x + t2 z is visible to the other scan warps
            if(SPEC) scan_now = spec;
          }
        }
        else if(SPEC && NSW > 1 && zw != 0 && warp > SW)
        {
          asm volatile("bar.sync 2, %0;" ::"n"(32 * (NSW > 1 ? NSW : 1)) : "memory");
          scan_now = dec[4] != 0;
        }
        if(SPEC)
        {
          xv = xs2;
          excl = sc.p; // the constraint being added is still INACTIVE in stat (or being written by warp 0)
        }
        // ---- warp 0: r = R^-1 d1, step length, the step itself and, when a constraint is added, the
        //      bookkeeping of the add
        if(warp == 0)
        {
          back_substitution();
          PH_MARK(5);
          double t1;
          len_t1(t1, l, !pre);
          if(W > 1 && zw != 0)
          {
            asm volatile("bar.sync 1, 64;" ::: "memory");
            t2 = decd[1];
            nz = decd[2];
            zpos = dec[3] != 0;
          }
          PH_MARK(7);
          double t;
          bool primal = true;
          add = true;
          int brk = -1;
          if(pre)
            t = zpos ? t2 : 0.0; // exact step onto the constraint (src/GoldfarbIdnaniSolver.cpp:307-322)
          else
          {
            t = t2 < t1 ? t2 : t1; // std::min(t1, t2)
            if(t >= big)
              brk = TS_INFEASIBLE;
            else if(t2 >= big)
              primal = add = false; // dual-only step, then drop
            else
              add = t == t2; // full step -> add ; partial step -> drop
          }
          int nvalid = 0;
          if(brk < 0)
          {
            take_step(t, nz, primal);
            PH_MARK(8);
            if(add)
            {
              // DualSolver::addConstraint bookkeeping (src/DualSolver.cpp:231-235)
              if(lane == 0)
              {
                alist[q] = sc.p;
                stat[sc.p] = (signed char)sc.st;
              }
              nvalid = want_next ? 1 : 0;
            }
          }
          if(lane == 0)
          {
            dec[0] = add;
            dec[1] = l;
            dec[2] = brk;
            dec[5] = nvalid;
            decd[0] = f;
          }
        }
        if(warp == CW)
        {
          givens_recurrence();
          PH_MARK(6);
        }
        if(!SPEC)
        {
          // every thread learns the decision; the scan (if wanted) then runs on the final x
          sync();
          PH_MARK(11);
          add = dec[0] != 0;
          l = dec[1];
          f = decd[0];
          if(dec[2] >= 0)
          {
            status = dec[2];
            break;
          }
          scan_now = add && dec[5] != 0;
          if(add && SPLIT_CHAIN)
          {
            // the (c, s) pairs of the sweep, one rotation per thread (second division of makeGivens); published by the
            // barrier below
            if(tid >= q && tid <= n - 2) rotation_cs(tid);
          }
        }
      }
      // ---- the constraint scan (one call site): of the current x when this pass only selects or when every thread
      //      takes part (W <= 2), of the speculative x on the scan warps otherwise
      if(warp >= SW && scan_now)
      {
        scan_partial<NSW>(xv, excl, warp - SW);
        PH_MARK(1);
      }
      sync();
      PH_MARK(11);
      if(sel_only)
      {
        sc = scan_combine<NSW>(cx_sel);
        have_sel = true;
        continue;
      }
      if(SPEC)
      {
        add = dec[0] != 0;
        l = dec[1];
        f = decd[0];
        if(dec[2] >= 0)
        {
          status = dec[2];
          break;
        }
        // a full step reached exactly the point the scan warps tested: x + t z with t == t2 (same fma, same bits)
        have_sel = add && dec[5] != 0 && dec[4] != 0;
      }
      else
        have_sel = scan_now;
      if(add)
      {
        if(have_sel)
        {
          double cxn;
          sc = scan_combine<NSW>(cxn);
          cx_sel = cxn;
        }
        if(SPEC && SPLIT_CHAIN)
        {
          if(tid >= q && tid <= n - 2) rotation_cs(tid);
          sync();
        }
        // the normal of the next step's constraint, when it is known already, is fetched while the rotations are applied
        // (wide kernels only: +4.3 % at n = 128, -1 % at n = 50 where two more registers and the extra code cost more, profiles/r02k_ab_*.txt)
        const int pnext = W < 3 ? -1 : have_sel ? ((sc.st != ST_INACTIVE && sc.st < ST_LOWER_BOUND) ? sc.p : -1) : (next_pre >= 0 && next_pre < mc ? next_pre : -1);
        double cvn = 0.0;
        if(pnext >= 0 && tid < n) cvn = Cb[(long long)pnext * ldC + tid];
        add_constraint();
        if(pnext >= 0)
        {
          if(tid < n) cv[tid] = cvn; // published by the first barrier of compute_step
          cv_ready = true;
        }
        PH_MARK(9);
      }
      else
      {
        remove_constraint(l);
        PH_MARK(10);
      }
      if(!pre)
      {
        skip = !add;
        ++it;
      }
    }
    sync();
    PH_MARK(11);
    write_result(b, status, it);
    PH_MARK(12);
    PH_FLUSH;
  }

  __device__ void write_result(long long b, int status, int it)
  {
    double * xo = P.x + b * n;
    for(int i = tid; i < n; i += T) xo[i] = xs[i];
    if(P.u)
    {
      // DualSolver::multipliers (src/DualSolver.cpp:38-69): thread i looks for itself in the active list
      double * uo = P.u + b * m;
      for(int i = tid; i < m; i += T)
      {
        int s = stat[i];
        double v = 0.0;
        if(s != ST_INACTIVE)
        {
          for(int k = 0; k < q; ++k)
            if(alist[k] == i) v = (s == ST_UPPER || s == ST_UPPER_BOUND) ? us[k] : -us[k];
        }
        uo[i] = v;
      }
    }
    if(P.active_set)
    {
      signed char * ao = P.active_set + b * m;
      for(int i = tid; i < m; i += T) ao[i] = stat[i];
    }
    if(P.active_list)
    {
      int * lo = P.active_list + b * n;
      for(int k = tid; k < n; k += T) lo[k] = k < q ? alist[k] : -1;
    }
    if(tid == 0)
    {
      if(P.f) P.f[b] = f;
      if(P.iterations) P.iterations[b] = it;
      if(P.status) P.status[b] = status;
      if(P.n_active) P.n_active[b] = q;
    }
  }

  __device__ void write_failure(long long b, int status)
  {
    double * xo = P.x + b * n;
    for(int i = tid; i < n; i += T) xo[i] = 0.0;
    if(P.u)
      for(int i = tid; i < m; i += T) P.u[b * m + i] = 0.0;
    if(P.active_set)
      for(int i = tid; i < m; i += T) P.active_set[b * m + i] = ST_INACTIVE;
    if(P.active_list)
      for(int k = tid; k < n; k += T) P.active_list[b * n + k] = -1;
    if(tid == 0)
    {
      if(P.f) P.f[b] = 0.0;
      if(P.iterations) P.iterations[b] = 0;
      if(P.status) P.status[b] = status;
      if(P.n_active) P.n_active[b] = 0;
    }
  }
};

// Persistent kernel: grid = resident CTAs of the whole GPU; every CTA pulls the next problem index
// from an atomic ticket counter, which absorbs the divergent iteration counts across QPs.
// MINB > 0: a second instantiation compiled for that many resident CTAs per SM (register cap). W = 3 only: 218 registers
// leave 2 warps per scheduler, i.e. 2 CTAs per SM whatever the shared memory allows; capped at 168 registers the kernel
// runs 3 QPs per SM where shared memory permits (n <= 77) and is 29 % faster there, 1.5 % slower where it does not
// (profiles/r4a_ab_w3cap.txt) — the host picks by occupancy (capi.cu).
template<int W, bool STAGE_C, bool WARM = false, int MINB = 0>
__global__ void __launch_bounds__(32 * W, (MINB ? MINB : (WARM ? (W == 1 ? JRLQP_WARM_MINB1 : (W == 2 ? JRLQP_WARM_MINB2 : 1)) : (W == 1 ? JRLQP_MINB1 : (W == 2 ? 6 : (W == 3 ? JRLQP_MINB3 : 1)))))) gi_dense_cta_kernel(const GiParams p)
{
  extern __shared__ __align__(16) double smem[];
  GiCta<W, STAGE_C, WARM> cta(p, smem);
  unsigned long long * ticket = reinterpret_cast<unsigned long long *>(smem + p.off_scr + 12);
  // the constraint scan reads x (and the speculative x kept in cv) in whole 16-byte pairs: the padding stays zero
  if((int)threadIdx.x >= p.n) cta.xs[threadIdx.x] = cta.cv[threadIdx.x] = 0.0;
  int ct_slot = -1;
  if(!STAGE_C && (W >= 3 || JRLQP_CT_ALLW) && p.ct != nullptr && p.mc > 0)
  {
    // claim a slice for the transposed copy of C (as many slices as CTAs of these kernels can be resident)
    if(threadIdx.x == 0)
    {
      int i = (int)(blockIdx.x % (unsigned)p.ct_slots);
      while(atomicCAS(p.ct_busy + i, 0, 1) != 0) i = i + 1 == p.ct_slots ? 0 : i + 1;
      __threadfence();
      *ticket = (unsigned long long)i;
    }
    cta.sync();
    ct_slot = (int)*ticket;
    cta.Ct = p.ct + (long long)ct_slot * p.ct_stride;
    cta.sync();
  }
#if JRLQP_BULK_PREFETCH
  // Ticket look-ahead: the index of the NEXT problem is drawn before the current one is solved, and its matrices are
  // pulled towards L2 by the TMA engine (one bulk prefetch per array: no register, no load instruction per line), so
  // that the staging loads of the next init() find them there instead of waiting on HBM.
  if(threadIdx.x == 0) *ticket = atomicAdd(p.counter, 1ull);
  cta.sync();
  unsigned long long b = *ticket;
  cta.sync();
  while(b < (unsigned long long)p.batch)
  {
    if(threadIdx.x == 0)
    {
      const unsigned long long nb = atomicAdd(p.counter, 1ull);
      *ticket = nb;
      if(nb < (unsigned long long)p.batch)
      {
        auto bulk = [](const double * ptr, long long doubles)
        {
          const unsigned bytes = (unsigned)(doubles * 8);
          if((reinterpret_cast<unsigned long long>(ptr) & 15ull) == 0ull && (bytes & 15u) == 0u && bytes > 0u)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ptr), "r"(bytes) : "memory");
        };
        if(p.sG != 0) bulk(p.G + nb * p.sG, (long long)(p.n - 1) * p.ldg + p.n);
        if(p.sC != 0 && p.mc > 0) bulk(p.C + nb * p.sC, (long long)(p.mc - 1) * p.ldc + p.n);
      }
    }
    cta.solve((long long)b);
    cta.sync();
    b = *ticket;
    cta.sync();
  }
#else
  for(;;)
  {
    if(threadIdx.x == 0) *ticket = atomicAdd(p.counter, 1ull);
    cta.sync();
    const unsigned long long b = *ticket;
    if(b >= (unsigned long long)p.batch) break;
    cta.solve((long long)b);
    cta.sync();
  }
#endif
  if(ct_slot >= 0 && threadIdx.x == 0)
  {
    __threadfence();
    atomicExch(p.ct_busy + ct_slot, 0);
  }
}

} // namespace jrlqp
