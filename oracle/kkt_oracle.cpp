// TEST INFRASTRUCTURE — NOT PRODUCT CODE. CPU restatement of the reference's KKT checker
//   jrl::qp::test::testKKT / testKKTStationarity / testKKTFeasibility / checkKKTConstraint
//   (src/test/kkt.cpp:14-195, include/jrl-qp/test/kkt.h:83-84: tau_p = tau_d = 1e-6)
// plus the planted-solution comparison the reference's tests make with Eigen's isApprox
//   (tests/GoldfarbIdnaniSolverTest.cpp:94-97: x.isApprox(pb.x, 1e-6)  <=>
//    |x - x*|^2 <= prec^2 min(|x|^2, |x*|^2)),
// batched, with ONE fixed floating-point order so that the GPU verifier (jrl-qp_b200/csrc/kkt.cu) can be
// compared with it bit for bit:
//   dL_i  = ((dot4_j(G(i,j), x_j) + a_i) [+ u_{mc+i}]) + dot4_c(C(i,c), u_c)      (src/test/kkt.cpp:121-134)
//   cx_c  = dot4_i(C(i,c), x_i)                                                     (:169-173)
//   norms = max |.| (exact whatever the order), squared norms in the dot32 order.
// The checker is tolerance based in the reference; only this restatement's own GPU twin is held to bits.
#include "gi_oracle.hpp"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <thread>
#include <vector>

using gi_oracle::dot32;
using gi_oracle::dot4;

static inline double max_nan(double a, double b)
{
  return a != a ? a : (b != b ? b : (a < b ? b : a));
}

namespace
{
// src/test/kkt.cpp:14-23
bool checkKKTConstraint(double cx, double bl, double bu, double u, double tau_x, double tau_u)
{
  double li = cx - bl;
  double ui = cx - bu;
  bool b1 = std::abs(li) <= tau_x && u <= -tau_u;
  bool b2 = li >= -tau_x && ui <= tau_x && std::abs(u) <= tau_u;
  bool b3 = std::abs(ui) <= tau_x && u >= tau_u;
  return b1 || b2 || b3;
}
} // namespace

extern "C"
{

/** flags[b]: bit 0 stationarity (src/test/kkt.cpp:105-137), bit 1 feasibility (:149-183), bit 2 planted
 * solution (only when x_ref is given). resid[b][4] (nullable): |dL|_inf, tau_u, tau_x, |x - x_ref|^2.
 * C is n x mc column-major (one normal per column, "transposedC" in the reference's vocabulary). */
int kkt_oracle_check_batch(int n,
                           int mc,
                           int nb,
                           long batch,
                           const double * G,
                           long sG,
                           int ldg,
                           const double * a,
                           long sa,
                           const double * C,
                           long sC,
                           int ldc,
                           const double * bl,
                           long sbl,
                           const double * bu,
                           long sbu,
                           const double * xl,
                           long sxl,
                           const double * xu,
                           long sxu,
                           const double * x,
                           const double * u,
                           const double * x_ref,
                           double tau_p,
                           double tau_d,
                           double prec,
                           int * flags,
                           double * resid,
                           int nthreads)
{
  const int m = mc + nb;
  if(nthreads < 1) nthreads = 1;
  std::atomic<long> next(0);
  std::atomic<long> nfail(0);
  auto worker = [&]()
  {
    std::vector<double> diff(static_cast<size_t>(n));
    for(;;)
    {
      const long b = next.fetch_add(1);
      if(b >= batch) break;
      const double * Gb = G + b * sG;
      const double * ab = a + b * sa;
      const double * Cb = mc ? C + b * sC : nullptr;
      const double * xb = x + b * n;
      const double * ub = u + b * m;
      double nx = 0, nu = 0;
      // (NaN-propagating maxima: a NaN solution must fail the `<= tau` tests, as it does in the reference)
      for(int i = 0; i < n; ++i) nx = max_nan(nx, std::abs(xb[i]));
      for(int i = 0; i < m; ++i) nu = max_nan(nu, std::abs(ub[i]));
      const double tau_x = tau_p * (1 + nx);
      const double tau_u = tau_d * (1 + nu);
      // stationarity
      double ndL = 0;
      for(int i = 0; i < n; ++i)
      {
        double t = dot4(n, Gb + i, ldg, xb, 1) + ab[i];
        if(nb) t = t + ub[mc + i];
        if(mc) t = t + dot4(mc, Cb + i, ldc, ub, 1);
        ndL = max_nan(ndL, std::abs(t));
      }
      int fl = 0;
      if(ndL <= tau_u) fl |= 1;
      // feasibility
      bool ok = true;
      for(int c = 0; c < mc && ok; ++c)
      {
        const double cx = dot4(n, Cb + static_cast<long>(c) * ldc, 1, xb, 1);
        ok = checkKKTConstraint(cx, bl[b * sbl + c], bu[b * sbu + c], ub[c], tau_x, tau_u);
      }
      for(int i = 0; i < nb && ok; ++i) ok = checkKKTConstraint(xb[i], xl[b * sxl + i], xu[b * sxu + i], ub[mc + i], tau_x, tau_u);
      if(ok) fl |= 2;
      double d2 = 0;
      if(x_ref)
      {
        const double * xr = x_ref + b * n;
        for(int i = 0; i < n; ++i) diff[i] = xb[i] - xr[i];
        d2 = dot32(n, diff.data(), diff.data());
        const double n1 = dot32(n, xb, xb), n2 = dot32(n, xr, xr);
        if(d2 <= (prec * prec) * std::min(n1, n2)) fl |= 4;
      }
      flags[b] = fl;
      if(resid)
      {
        resid[4 * b] = ndL;
        resid[4 * b + 1] = tau_u;
        resid[4 * b + 2] = tau_x;
        resid[4 * b + 3] = d2;
      }
      const int want = x_ref ? 7 : 3;
      if(fl != want) nfail.fetch_add(1);
    }
  };
  std::vector<std::thread> th;
  for(int t = 1; t < nthreads; ++t) th.emplace_back(worker);
  worker();
  for(auto & t : th) t.join();
  return static_cast<int>(std::min<long>(nfail.load(), 0x7fffffff));
}

} // extern "C"
