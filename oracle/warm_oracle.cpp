// TEST INFRASTRUCTURE — NOT PRODUCT CODE. See gi_oracle.hpp for scope.
//
// CPU restatement of the reference's warm-start capable solver,
// experimental::GoldfarbIdnaniSolver (src/experimental/GoldfarbIdnaniSolver.cpp):
//   :21-64    solve(..., as)                 -> GIOracle::solveExperimental
//   :66-111   init_ (incl. the loop dropping constraints activated with u < 0) -> initExperimental
//   :306-381  processInitialActiveSet        -> processInitialActiveSet
//   :383-459  initializeComputationData      -> initializeComputationData
//   :461-486  initializePrimalDualPoints     -> initializePrimalDualPoints
// The iteration hooks (:113-304) are the same as the stable solver's and are shared (gi_oracle.cpp).
//
// Eigen primitives restated (canonical order, same family as gi_oracle.hpp):
//   B = L^-1 N                       B(r,k) = (N(r,k) - dot4_{j<r}(L(r,j), B(j,k))) / L(r,r)
//   householder_qr_inplace           unblocked, column by column (Eigen switches to a blocked update
//                                    beyond 48 columns; one canonical order is kept here for all q):
//       makeHouseholder              tailSq = dot32(tail, tail); if tailSq <= DBL_MIN: tau = 0, beta = c0,
//                                    essential = 0; else beta = -sign(c0) sqrt(fma(c0, c0, tailSq)),
//                                    essential = tail / (c0 - beta), tau = (beta - c0) / beta
//       applyHouseholderOnTheLeft    tmp_j = dot4(essential, bottom_j) + top_j; top_j = fma(-tau, tmp_j, top_j);
//                                    bottom(i,j) = fma(-(tau essential_i), tmp_j, bottom(i,j))
//   HouseholderSequence::applyThisOnTheRight (J = J Q), k ascending:
//                                    tmp_i = dot4_j(J(i,k+1+j), essential_j) + J(i,k); J(i,k) = fma(-tau, tmp_i, J(i,k));
//                                    J(i,k+1+j) = fma(-(tau tmp_i), essential_j, J(i,k+1+j))
//   alpha = J^T a                    dot4 per entry;  beta = R^-T b_act, u = R^-1 (alpha1 + beta): column-oriented
//                                    substitution with true division;  x_i = dot4_{j<q}(J(i,j), beta_j)
//                                    - dot4_{j>=q}(J(i,j), alpha_j);  f = dot32(beta, fma(0.5, beta, alpha1))
//                                    - 0.5 dot32(alpha2, alpha2)
#include "gi_oracle.hpp"

#include <algorithm>
#include <cassert>
#include <cfloat>
#include <cmath>

namespace gi_oracle
{

TerminationStatus GIOracle::solveExperimental(double * G,
                                              int ldg,
                                              const double * a,
                                              const double * C,
                                              int ldc,
                                              const double * bl,
                                              const double * bu,
                                              const double * xl,
                                              const double * xu,
                                              const int8_t * as)
{
  G_ = G;
  ldg_ = ldg;
  a_ = a;
  C_ = C;
  ldc_ = ldc;
  bl_ = bl;
  bu_ = bu;
  xl_ = xl;
  xu_ = xu;
  flops_ = 0;
  margin_ = INFINITY;
  trace_.clear();
  const int m = A_.nbAll();
  // :58-61  pb_.as = as.empty() ? A_.activationStatus() : as   (only when warm start is on)
  if(opt_.warmStart)
  {
    if(as)
    {
      asIn_.resize(static_cast<size_t>(m));
      for(int i = 0; i < m; ++i) asIn_[static_cast<size_t>(i)] = static_cast<ActivationStatus>(as[i]);
    }
    else
      asIn_ = A_.activationStatus();
  }
  // DualSolver::init (src/DualSolver.cpp:200-210)
  needExpand_ = true;
  if(!opt_.warmStart) A_.reset();
  it_ = 0;
  TerminationStatus st = initExperimental();
  if(st != SUCCESS) return st;
  return mainLoop();
}

TerminationStatus GIOracle::initExperimental()
{
  const int n = n_;
  if(static_cast<int>(bact_.size()) < n)
  {
    bact_.assign(static_cast<size_t>(n), 0.0);
    hcoef_.assign(static_cast<size_t>(n), 0.0);
    alpha_.assign(static_cast<size_t>(n), 0.0);
  }
  TerminationStatus st = processInitialActiveSet();
  if(st != SUCCESS) return st;

  // :75-77 Cholesky, same canonical order as the stable solver (gi_oracle.cpp init())
  double * G = G_;
  const std::ptrdiff_t ld = ldg_;
  double * w = w_.data();
  for(int k = 0; k < n; ++k)
  {
    for(int i = k; i < n; ++i) w[i] = G[i + k * ld] - dot4(k, G + i, ld, G + k, ld);
    if(w[k] <= 0.0) return NON_POS_HESSIAN;
    double lkk = std::sqrt(w[k]);
    G[k + k * ld] = lkk;
    for(int i = k + 1; i < n; ++i) G[i + k * ld] = w[i] / lkk;
  }

  initializeComputationData();
  initializePrimalDualPoints();

  // :83-108 constraints activated with a negative multiplier are dropped, most negative first
  for(;;)
  {
    int q = A_.nbActiveCstr();
    double * u = u_.data();
    double umin = -1e-14;
    int lmin = -1;
    for(int l = 0; l < q; ++l)
    {
      int i = A_[l];
      if(instrument_) noteMargin(u[l], umin);
      if(u[l] < umin && A_.activationStatus(i) != FIXED && A_.activationStatus(i) != EQUALITY)
      {
        umin = u[l];
        lmin = l;
      }
    }
    if(lmin < 0) break;
    ++it_;
    for(int k = lmin; k < q - 1; ++k) bact_[static_cast<size_t>(k)] = bact_[static_cast<size_t>(k + 1)];
    A_.deactivate(lmin);
    removeConstraintCore(lmin);
    initializePrimalDualPoints();
  }
  return SUCCESS;
}

TerminationStatus GIOracle::processInitialActiveSet()
{
  // src/experimental/GoldfarbIdnaniSolver.cpp:306-381
  A_.reset();
  const int mc = A_.nbCstr(), nb = A_.nbBnd();
  const bool useAs = !asIn_.empty() && opt_.warmStart;
  for(int i = 0; i < nb; ++i)
  {
    int bi = mc + i;
    if(xl_[i] == xu_[i])
      A_.activate(bi, FIXED);
    else if(useAs && asIn_[static_cast<size_t>(bi)] != INACTIVE)
    {
      ActivationStatus s = asIn_[static_cast<size_t>(bi)];
      if(s == FIXED)
        ; // ignored: the bounds are not equal
      else if((s == LOWER_BOUND && xl_[i] < -opt_.bigBnd) || (s == UPPER_BOUND && xu_[i] > +opt_.bigBnd))
        ; // ignored: infinite bound
      else if(s != LOWER_BOUND && s != UPPER_BOUND)
        ; // not a bound status: the reference asserts (s > EQUALITY) and is undefined otherwise; ignored here
      else
        A_.activate(bi, s);
    }
  }
  for(int i = 0; i < mc; ++i)
  {
    if(bl_[i] == bu_[i])
      A_.activate(i, EQUALITY);
    else if(useAs && asIn_[static_cast<size_t>(i)] != INACTIVE)
    {
      ActivationStatus s = asIn_[static_cast<size_t>(i)];
      if(s == FIXED)
        ; // ignored
      else if((s == LOWER && bl_[i] < -opt_.bigBnd) || (s == UPPER && bu_[i] > +opt_.bigBnd))
        ; // ignored: infinite bound
      else if(s > EQUALITY)
        ; // a bound status on a general constraint: the reference asserts (s <= EQUALITY); ignored here
      else
        A_.activate(i, s); // NOTE: an EQUALITY guess on bl != bu is honoured, as in the reference (:352-368)
    }
  }
  if(A_.nbActiveCstr() > n_)
  {
    if(A_.nbActiveEquality() + A_.nbFixedVariable() > n_) return OVERCONSTRAINED_PROBLEM;
    auto isEqualityOrFixed = [this](int i)
    {
      ActivationStatus a = A_.activationStatus(A_[i]);
      return a == EQUALITY || a == FIXED;
    };
    int i = A_.nbActiveCstr();
    while(A_.nbActiveCstr() > n_)
    {
      --i;
      while(isEqualityOrFixed(i)) --i;
      A_.deactivate(i);
    }
  }
  return SUCCESS;
}

void GIOracle::initializeComputationData()
{
  // src/experimental/GoldfarbIdnaniSolver.cpp:383-459
  const int n = n_;
  const int mc = A_.nbCstr();
  const int q = A_.nbActiveCstr();
  const double * L = G_;
  const std::ptrdiff_t ld = ldg_;
  double * N = R_.data(); // n x q, ld n: the active normals, then B, then R (upper) + Householder vectors
  double * bact = bact_.data();
  for(int k = 0; k < q; ++k)
  {
    double * Nk = N + static_cast<size_t>(k) * n;
    int ci = A_[k];
    switch(A_.activationStatus(ci))
    {
      case LOWER:
      case EQUALITY:
        for(int i = 0; i < n; ++i) Nk[i] = C_[static_cast<size_t>(ci) * ldc_ + i];
        bact[k] = bl_[ci];
        break;
      case UPPER:
        for(int i = 0; i < n; ++i) Nk[i] = -C_[static_cast<size_t>(ci) * ldc_ + i];
        bact[k] = -bu_[ci];
        break;
      case LOWER_BOUND:
      case FIXED:
        for(int i = 0; i < n; ++i) Nk[i] = 0;
        Nk[ci - mc] = 1;
        bact[k] = xl_[ci - mc];
        break;
      case UPPER_BOUND:
        for(int i = 0; i < n; ++i) Nk[i] = 0;
        Nk[ci - mc] = -1;
        bact[k] = -xu_[ci - mc];
        break;
      default:
        break;
    }
  }

  // J = L^-T (same canonical order as the stable solver)
  double * J = J_.data();
  std::fill(J_.begin(), J_.end(), 0.0);
  double * rinv = d_.data();
  for(int i = 0; i < n; ++i) rinv[i] = 1.0 / L[i + i * ld];
  for(int j = 0; j < n; ++j)
  {
    double * Jj = J + static_cast<size_t>(j) * n;
    Jj[j] = rinv[j];
    for(int i = j - 1; i >= 0; --i)
    {
      double s = dot4(j - i, L + (i + 1) + i * ld, 1, Jj + i + 1, 1);
      Jj[i] = (-s) * rinv[i];
    }
  }

  // B = L^-1 N
  for(int k = 0; k < q; ++k)
  {
    double * Bk = N + static_cast<size_t>(k) * n;
    for(int r = 0; r < n; ++r) Bk[r] = (Bk[r] - dot4(r, L + r, ld, Bk, 1)) / L[r + r * ld];
  }

  // Householder QR of B in place
  double * h = hcoef_.data();
  for(int k = 0; k < q; ++k)
  {
    double * Bk = N + static_cast<size_t>(k) * n;
    const int len = n - k - 1; // size of the essential part
    double c0 = Bk[k];
    double tailSq = dot32(len, Bk + k + 1, Bk + k + 1);
    double tau, beta;
    if(tailSq <= DBL_MIN)
    {
      tau = 0;
      beta = c0;
      for(int i = 0; i < len; ++i) Bk[k + 1 + i] = 0;
    }
    else
    {
      beta = std::sqrt(std::fma(c0, c0, tailSq));
      if(c0 >= 0) beta = -beta;
      double den = c0 - beta;
      for(int i = 0; i < len; ++i) Bk[k + 1 + i] = Bk[k + 1 + i] / den;
      tau = (beta - c0) / beta;
    }
    Bk[k] = beta;
    h[k] = tau;
    // apply H_k to the remaining columns (Eigen skips the update when tau == 0, and scales by
    // 1 - tau when a single row is left)
    for(int j = k + 1; j < q; ++j)
    {
      double * Bj = N + static_cast<size_t>(j) * n;
      if(len == 0)
        Bj[k] = Bj[k] * (1.0 - tau);
      else if(tau != 0)
      {
        double tmp = dot4(len, Bk + k + 1, 1, Bj + k + 1, 1) + Bj[k];
        Bj[k] = std::fma(-tau, tmp, Bj[k]);
        for(int i = 0; i < len; ++i) Bj[k + 1 + i] = std::fma(-(tau * Bk[k + 1 + i]), tmp, Bj[k + 1 + i]);
      }
    }
  }

  // J = J Q, Q = H_0 H_1 ... H_{q-1}
  for(int k = 0; k < q; ++k)
  {
    const double * ess = N + static_cast<size_t>(k) * n + k + 1;
    const int len = n - k - 1;
    const double tau = h[k];
    double * Jk = J + static_cast<size_t>(k) * n;
    if(len == 0)
    {
      for(int i = 0; i < n; ++i) Jk[i] = Jk[i] * (1.0 - tau);
    }
    else if(tau != 0)
    {
      for(int i = 0; i < n; ++i)
      {
        double tmp = dot4(len, Jk + n + i, n, ess, 1) + Jk[i];
        Jk[i] = std::fma(-tau, tmp, Jk[i]);
        double tt = tau * tmp;
        for(int j = 0; j < len; ++j) Jk[static_cast<size_t>(j + 1) * n + i] = std::fma(-tt, ess[j], Jk[static_cast<size_t>(j + 1) * n + i]);
      }
    }
  }
  if(instrument_)
  {
    double dn = n, dq = q;
    flops_ += dn * dn * dn / 3.0 + dn * dn * dn / 3.0 + dq * dn * dn + 2.0 * dn * dq * dq + 4.0 * dn * dn * dq;
  }
}

void GIOracle::initializePrimalDualPoints()
{
  // src/experimental/GoldfarbIdnaniSolver.cpp:461-486
  const int n = n_;
  const int q = A_.nbActiveCstr();
  const double * J = J_.data();
  const double * R = R_.data();
  double * alpha = alpha_.data();
  double * beta = r_.data(); // work_r_
  double * x = x_.data();
  double * u = u_.data();
  const double * bact = bact_.data();
  double * w = w_.data();

  for(int j = 0; j < n; ++j) alpha[j] = dot4(n, J + static_cast<size_t>(j) * n, 1, a_, 1);
  // beta = R^-T b_act (R^T lower triangular): column-oriented forward substitution
  for(int k = 0; k < q; ++k) w[k] = bact[k];
  for(int k = 0; k < q; ++k)
  {
    double bk = w[k] / R[k + static_cast<size_t>(k) * n];
    beta[k] = bk;
    for(int i = k + 1; i < q; ++i) w[i] = std::fma(-bk, R[k + static_cast<size_t>(i) * n], w[i]);
  }
  // x = J1 beta - J2 alpha2
  for(int i = 0; i < n; ++i) x[i] = dot4(q, J + i, n, beta, 1) - dot4(n - q, J + static_cast<size_t>(q) * n + i, n, alpha + q, 1);
  // u = R^-1 (alpha1 + beta)
  for(int k = 0; k < q; ++k) w[k] = alpha[k] + beta[k];
  for(int k = q - 1; k >= 0; --k)
  {
    double uk = w[k] / R[k + static_cast<size_t>(k) * n];
    u[k] = uk;
    for(int j = 0; j < k; ++j) w[j] = std::fma(-uk, R[j + static_cast<size_t>(k) * n], w[j]);
  }
  // f = beta.(0.5 beta + alpha1) - 0.5 |alpha2|^2
  for(int k = 0; k < q; ++k) w[k] = std::fma(0.5, beta[k], alpha[k]);
  f_ = dot32(q, beta, w) - 0.5 * dot32(n - q, alpha + q, alpha + q);
  if(instrument_)
  {
    double dn = n, dq = q;
    flops_ += 2.0 * dn * dn + 2.0 * dq * dq + 2.0 * dn * dn + 4.0 * dn;
  }
}

} // namespace gi_oracle
