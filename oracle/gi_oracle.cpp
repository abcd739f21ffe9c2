// TEST INFRASTRUCTURE — NOT PRODUCT CODE. See gi_oracle.hpp for the scope,
// the reference files restated and the canonical arithmetic order.
#include "gi_oracle.hpp"

#include <algorithm>
#include <cassert>
#include <cmath>
#include <limits>

namespace gi_oracle
{

// ---------------------------------------------------------------------------
// canonical primitives
// ---------------------------------------------------------------------------

double dot4(int len, const double * a, std::ptrdiff_t sa, const double * b, std::ptrdiff_t sb)
{
  double c0 = 0, c1 = 0, c2 = 0, c3 = 0;
  int k = 0;
  for(; k + 3 < len; k += 4)
  {
    c0 = std::fma(a[(k + 0) * sa], b[(k + 0) * sb], c0);
    c1 = std::fma(a[(k + 1) * sa], b[(k + 1) * sb], c1);
    c2 = std::fma(a[(k + 2) * sa], b[(k + 2) * sb], c2);
    c3 = std::fma(a[(k + 3) * sa], b[(k + 3) * sb], c3);
  }
  if(k < len) c0 = std::fma(a[k * sa], b[k * sb], c0);
  if(k + 1 < len) c1 = std::fma(a[(k + 1) * sa], b[(k + 1) * sb], c1);
  if(k + 2 < len) c2 = std::fma(a[(k + 2) * sa], b[(k + 2) * sb], c2);
  return (c0 + c1) + (c2 + c3);
}

double dot32(int len, const double * a, const double * b)
{
  double acc[32];
  for(int l = 0; l < 32; ++l) acc[l] = 0;
  for(int k = 0; k < len; ++k) acc[k & 31] = std::fma(a[k], b[k], acc[k & 31]);
  for(int off = 16; off >= 1; off >>= 1)
  {
    double nxt[32];
    for(int l = 0; l < 32; ++l) nxt[l] = acc[l] + acc[l ^ off];
    for(int l = 0; l < 32; ++l) acc[l] = nxt[l];
  }
  return acc[0];
}

void makeGivens(double p, double q, double & c, double & s, double & r)
{
  if(q == 0.0)
  {
    c = p < 0.0 ? -1.0 : 1.0;
    s = 0.0;
    r = std::abs(p);
  }
  else if(p == 0.0)
  {
    c = 0.0;
    s = q < 0.0 ? 1.0 : -1.0;
    r = std::abs(q);
  }
  else if(std::abs(p) > std::abs(q))
  {
    double t = q / p;
    double u = std::sqrt(std::fma(t, t, 1.0));
    if(p < 0.0) u = -u;
    c = 1.0 / u;
    s = -t * c;
    r = p * u;
  }
  else
  {
    double t = p / q;
    double u = std::sqrt(std::fma(t, t, 1.0));
    if(q < 0.0) u = -u;
    s = -1.0 / u;
    c = -t * s;
    r = q * u;
  }
}

// Rotation of two strided vectors: x' = c x - s y ; y' = s x + c y
// (Eigen applyOnTheRight(i, i+1, G) on columns / applyOnTheLeft(i, i+1, G^T) on rows).
static inline void rotate(int len, double * x, std::ptrdiff_t sx, double * y, std::ptrdiff_t sy, double c, double s)
{
  for(int k = 0; k < len; ++k)
  {
    double xi = x[k * sx];
    double yi = y[k * sy];
    x[k * sx] = std::fma(c, xi, -(s * yi));
    y[k * sy] = std::fma(c, yi, s * xi);
  }
}

// ---------------------------------------------------------------------------
// ActiveSet (src/internal/ActiveSet.cpp:30-168)
// ---------------------------------------------------------------------------

void ActiveSet::resize(int nCstr, int nBnd)
{
  assert(nCstr >= 0 && nBnd >= 0);
  size_t nTot = static_cast<size_t>(nCstr) + static_cast<size_t>(nBnd);
  status_.resize(nTot);
  activeSet_.reserve(nTot);
  nbCstr_ = nCstr;
  nbBnd_ = nBnd;
  reset();
}

void ActiveSet::reset()
{
  std::fill(status_.begin(), status_.end(), INACTIVE);
  activeSet_.clear();
  me_ = mi_ = ml_ = mu_ = mb_ = mbl_ = mbu_ = mbe_ = 0;
}

void ActiveSet::count(ActivationStatus s, int delta)
{
  switch(s)
  {
    case LOWER:
      mi_ += delta;
      ml_ += delta;
      break;
    case UPPER:
      mi_ += delta;
      mu_ += delta;
      break;
    case EQUALITY:
      me_ += delta;
      break;
    case LOWER_BOUND:
      mb_ += delta;
      mbl_ += delta;
      break;
    case UPPER_BOUND:
      mb_ += delta;
      mbu_ += delta;
      break;
    case FIXED:
      mb_ += delta;
      mbe_ += delta;
      break;
    default:
      assert(false);
  }
}

void ActiveSet::activate(int cstrIdx, ActivationStatus status)
{
  assert(cstrIdx < nbCstr_ + nbBnd_);
  assert(status_[static_cast<size_t>(cstrIdx)] == INACTIVE);
  assert(status != INACTIVE);
  activeSet_.push_back(cstrIdx);
  status_[static_cast<size_t>(cstrIdx)] = status;
  count(status, +1);
}

void ActiveSet::deactivate(int activeIdx)
{
  int cstrIdx = activeSet_[static_cast<size_t>(activeIdx)];
  ActivationStatus status = status_[static_cast<size_t>(cstrIdx)];
  activeSet_.erase(activeSet_.begin() + activeIdx);
  status_[static_cast<size_t>(cstrIdx)] = INACTIVE;
  count(status, -1);
}

// ---------------------------------------------------------------------------
// GIOracle
// ---------------------------------------------------------------------------

GIOracle::GIOracle(int n, int mc, bool useBounds)
{
  resize(n, mc, useBounds);
}

void GIOracle::resize(int n, int mc, bool useBounds)
{
  // src/DualSolver.cpp:251-274 + src/GoldfarbIdnaniSolver.cpp:258-266
  int nb = useBounds ? n : 0;
  if(n != n_)
  {
    n_ = n;
    size_t nn = static_cast<size_t>(n) * static_cast<size_t>(n);
    x_.assign(static_cast<size_t>(n), 0.0);
    z_.assign(static_cast<size_t>(n), 0.0);
    d_.assign(static_cast<size_t>(n), 0.0);
    w_.assign(static_cast<size_t>(n), 0.0);
    acc_.assign(4 * static_cast<size_t>(n), 0.0);
    J_.assign(nn, 0.0);
    R_.assign(nn, 0.0);
  }
  if(mc + nb != A_.nbAll() || u_.empty())
  {
    u_.assign(static_cast<size_t>(mc + nb) + 1, 0.0);
    r_.assign(static_cast<size_t>(mc + nb) + 1, 0.0);
    uExp_.assign(static_cast<size_t>(mc + nb) + 1, 0.0);
  }
  if(mc != A_.nbCstr() || nb != A_.nbBnd()) A_.resize(mc, nb);
}

void GIOracle::noteMargin(double a, double b)
{
  double m = std::abs(a - b) / std::max(1.0, std::max(std::abs(a), std::abs(b)));
  if(m < margin_) margin_ = m;
}

TerminationStatus GIOracle::solve(double * G,
                                  int ldg,
                                  const double * a,
                                  const double * C,
                                  int ldc,
                                  const double * bl,
                                  const double * bu,
                                  const double * xl,
                                  const double * xu)
{
  // src/GoldfarbIdnaniSolver.cpp:18-54 — bind views (sizes were fixed by resize()).
  G_ = G;
  ldg_ = ldg;
  a_ = a;
  C_ = C;
  ldc_ = ldc;
  bl_ = bl;
  bu_ = bu;
  xl_ = xl;
  xu_ = xu;
  assert((xl != nullptr) == (A_.nbBnd() > 0));

  flops_ = 0;
  margin_ = std::numeric_limits<double>::infinity();
  trace_.clear();

  // src/DualSolver.cpp:200-210 init(): the stable solver always resets the active
  // set (src/GoldfarbIdnaniSolver.cpp:75), so warmStart is ignored here, as in the reference.
  needExpand_ = true;
  it_ = 0;
  if(!init()) return overconstrained_ ? OVERCONSTRAINED_PROBLEM : NON_POS_HESSIAN;
  return mainLoop();
}

// src/DualSolver.cpp:96-168 — the loop shared by the cold solver and the experimental (warm) one.
TerminationStatus GIOracle::mainLoop()
{
  const int n = n_;
  bool skipStep1 = false;
  Selected sc;
  double * x = x_.data();
  double * z = z_.data();
  double * u = u_.data();
  double * r = r_.data();

  for(; it_ < opt_.maxIter; ++it_)
  {
    int q = A_.nbActiveCstr();
    // Step 1
    if(!skipStep1)
    {
      sc = select();
      if(sc.st == INACTIVE) return SUCCESS;
      u[q] = 0;
    }
    // Step 2
    computeStep(sc);
    double t1, t2;
    int l;
    computeStepLength(sc, t1, t2, l);
    double t = std::min(t1, t2);
    if(instrument_) noteMargin(t1, t2);

    if(t >= opt_.bigBnd) return INFEASIBLE;

    if(t2 >= opt_.bigBnd)
    {
      for(int k = 0; k < q; ++k) u[k] = std::fma(-t, r[k], u[k]);
      u[q] += t;
      if(instrument_)
      {
        flops_ += 2.0 * q;
        trace_.push_back({it_, sc.p, sc.st, l, 2, t1, t2});
      }
      removeConstraint(l);
      skipStep1 = true;
    }
    else
    {
      for(int i = 0; i < n; ++i) x[i] = std::fma(t, z[i], x[i]);
      f_ += (t * normalDot(sc, z)) * (0.5 * t + u[q]);
      for(int k = 0; k < q; ++k) u[k] = std::fma(-t, r[k], u[k]);
      u[q] += t;
      if(instrument_) flops_ += 2.0 * n + 2.0 * q + 2.0 * n;
      if(t == t2)
      {
        if(instrument_) trace_.push_back({it_, sc.p, sc.st, l, 0, t1, t2});
        addConstraint(sc);
        skipStep1 = false;
      }
      else
      {
        if(instrument_) trace_.push_back({it_, sc.p, sc.st, l, 1, t1, t2});
        removeConstraint(l);
        skipStep1 = true;
      }
    }
  }
  return MAX_ITER_REACHED;
}

bool GIOracle::init()
{
  const int n = n_;
  double * G = G_;
  const std::ptrdiff_t ld = ldg_;
  double * w = w_.data();

  // --- In-place lower Cholesky (Eigen llt_inplace, src/GoldfarbIdnaniSolver.cpp:58-61).
  // Canonical order: left-looking by column k; for every row i >= k
  //   v_i = G(i,k) - dot4_{j<k}( L(i,j), L(k,j) );  L(k,k) = sqrt(v_k);  L(i,k) = v_i / L(k,k).
  for(int k = 0; k < n; ++k)
  {
    for(int i = k; i < n; ++i) w[i] = G[i + k * ld] - dot4(k, G + i, ld, G + k, ld);
    if(w[k] <= 0.0) return false; // Eigen: "if (x <= 0) return k" -> NON_POS_HESSIAN
    double lkk = std::sqrt(w[k]);
    G[k + k * ld] = lkk;
    for(int i = k + 1; i < n; ++i) G[i + k * ld] = w[i] / lkk;
  }

  // --- J = L^-T (src/GoldfarbIdnaniSolver.cpp:64-66). Upper triangular. Canonical order
  // (Eigen's trsm multiplies by the reciprocal of the diagonal): per column j,
  //   J(j,j) = 1/L(j,j);  J(i,j) = ( -dot4_{k=i+1..j}( L(k,i), J(k,j) ) ) * (1/L(i,i)),  i = j-1..0.
  double * J = J_.data();
  std::fill(J_.begin(), J_.end(), 0.0);
  double * rinv = d_.data(); // scratch
  for(int i = 0; i < n; ++i) rinv[i] = 1.0 / G[i + i * ld];
  for(int j = 0; j < n; ++j)
  {
    double * Jj = J + static_cast<size_t>(j) * n;
    Jj[j] = rinv[j];
    for(int i = j - 1; i >= 0; --i)
    {
      double s = dot4(j - i, G + (i + 1) + i * ld, 1, Jj + i + 1, 1);
      Jj[i] = (-s) * rinv[i];
    }
  }

  // --- x = -G^-1 a (src/GoldfarbIdnaniSolver.cpp:69-72). Canonical order: column-oriented
  // forward then backward substitution with true division by the diagonal.
  double * x = x_.data();
  for(int i = 0; i < n; ++i) w[i] = a_[i];
  for(int k = 0; k < n; ++k)
  {
    double yk = w[k] / G[k + k * ld];
    w[k] = yk;
    for(int i = k + 1; i < n; ++i) w[i] = std::fma(-yk, G[i + k * ld], w[i]);
  }
  for(int k = n - 1; k >= 0; --k)
  {
    double xk = w[k] / G[k + k * ld];
    x[k] = xk;
    for(int i = 0; i < k; ++i) w[i] = std::fma(-xk, G[k + i * ld], w[i]);
  }
  for(int i = 0; i < n; ++i) x[i] = -x[i];
  // f = 0.5 a.x (src/GoldfarbIdnaniSolver.cpp:73)
  f_ = 0.5 * dot32(n, a_, x);

  if(instrument_)
  {
    double dn = n;
    flops_ += dn * dn * dn / 3.0 + dn * dn * dn / 3.0 + 2.0 * dn * dn + 2.0 * dn;
  }

  A_.reset(); // src/GoldfarbIdnaniSolver.cpp:75
  initActiveSet(); // :79
  return !overconstrained_;
}

void GIOracle::initActiveSet()
{
  // src/GoldfarbIdnaniSolver.cpp:268-287. The reference pre-activates without looking at the count: with more than
  // nbVar equalities / fixed variables it writes past its workspaces (undefined behaviour). Here — and in the CUDA
  // kernels — the (nbVar + 1)-th pre-activation ends the solve with OVERCONSTRAINED_PROBLEM, the status the
  // experimental solver returns for the same input (src/experimental/GoldfarbIdnaniSolver.cpp:360-362).
  overconstrained_ = false;
  for(int i = 0; i < A_.nbCstr(); ++i)
  {
    if(bl_[i] == bu_[i])
    {
      if(A_.nbActiveCstr() >= n_)
      {
        overconstrained_ = true;
        return;
      }
      addInitialConstraint({i, EQUALITY});
    }
  }
  for(int i = 0; i < A_.nbBnd(); ++i)
  {
    if(xl_[i] == xu_[i])
    {
      if(A_.nbActiveCstr() >= n_)
      {
        overconstrained_ = true;
        return;
      }
      addInitialConstraint({A_.nbCstr() + i, FIXED});
    }
  }
}

void GIOracle::addInitialConstraint(Selected sc)
{
  // src/GoldfarbIdnaniSolver.cpp:295-338
  const int n = n_;
  int q = A_.nbActiveCstr();
  double * x = x_.data();
  double * z = z_.data();
  double * u = u_.data();
  double * r = r_.data();
  u[q] = 0;

  computeStep(sc);

  double t = 0;
  double znorm = std::sqrt(dot32(n, z, z));
  if(instrument_) noteMargin(znorm, 1e-14);
  if(znorm > 1e-14)
  {
    if(sc.st == EQUALITY)
    {
      const double * c = C_ + static_cast<size_t>(sc.p) * ldc_;
      t = (bl_[sc.p] - dot4(n, c, 1, x, 1)) / dot4(n, c, 1, z, 1);
    }
    else
    {
      int pb = sc.p - A_.nbCstr();
      t = (xl_[pb] - x[pb]) / z[pb];
    }
  }
  for(int i = 0; i < n; ++i) x[i] = std::fma(t, z[i], x[i]);
  f_ += (t * normalDot(sc, z)) * (0.5 * t + u[q]);
  for(int k = 0; k < q; ++k) u[k] = std::fma(-t, r[k], u[k]);
  u[q] += t;
  if(instrument_)
  {
    flops_ += (sc.st == EQUALITY ? 4.0 * n : 0.0) + 2.0 * n + q; // F_len
    flops_ += 2.0 * n + 2.0 * q + 2.0 * n; // F_update
    trace_.push_back({-1, sc.p, sc.st, 0, 3, 0.0, t});
  }
  addConstraint(sc);
}

GIOracle::Selected GIOracle::select()
{
  // src/GoldfarbIdnaniSolver.cpp:84-134. cx = dot4(C(:,i), x).
  const int n = n_;
  const double * x = x_.data();
  double smin = 0;
  Selected sel;
  const int mc = A_.nbCstr();
  for(int i = 0; i < mc; ++i)
  {
    if(!A_.isActive(i))
    {
      double cx = dot4(n, C_ + static_cast<size_t>(i) * ldc_, 1, x, 1);
      double sl = cx - bl_[i];
      if(instrument_) noteMargin(sl, smin);
      if(sl < smin)
      {
        smin = sl;
        sel = {i, LOWER};
      }
      else
      {
        double su = bu_[i] - cx;
        if(instrument_) noteMargin(su, smin);
        if(su < smin)
        {
          smin = su;
          sel = {i, UPPER};
        }
      }
    }
  }
  for(int i = 0; i < A_.nbBnd(); ++i)
  {
    if(!A_.isActiveBnd(i))
    {
      double sl = x[i] - xl_[i];
      if(instrument_) noteMargin(sl, smin);
      if(sl < smin)
      {
        smin = sl;
        sel = {mc + i, LOWER_BOUND};
      }
      else
      {
        double su = xu_[i] - x[i];
        if(instrument_) noteMargin(su, smin);
        if(su < smin)
        {
          smin = su;
          sel = {mc + i, UPPER_BOUND};
        }
      }
    }
  }
  if(instrument_) flops_ += 2.0 * n * mc;
  return sel;
}

void GIOracle::computeStep(Selected sc)
{
  // src/GoldfarbIdnaniSolver.cpp:136-148 + ConstraintNormal::preMultiplyByMt
  const int n = n_;
  const int q = A_.nbActiveCstr();
  double * d = d_.data();
  double * z = z_.data();
  double * r = r_.data();
  double * w = w_.data();
  const double * J = J_.data();
  const double * R = R_.data();

  // d = J^T n+ : d[j] = dot4_i( J(i,j), c[i] ), negated as a whole for UPPER;
  // +/- row of J for bounds.
  switch(sc.st)
  {
    case EQUALITY:
    case LOWER:
    case UPPER:
    {
      const double * c = C_ + static_cast<size_t>(sc.p) * ldc_;
      for(int j = 0; j < n; ++j) d[j] = dot4(n, J + static_cast<size_t>(j) * n, 1, c, 1);
      if(sc.st == UPPER)
        for(int j = 0; j < n; ++j) d[j] = -d[j];
      break;
    }
    case FIXED:
    case LOWER_BOUND:
    {
      int b = sc.p - A_.nbCstr();
      for(int j = 0; j < n; ++j) d[j] = J[b + static_cast<size_t>(j) * n];
      break;
    }
    case UPPER_BOUND:
    {
      int b = sc.p - A_.nbCstr();
      for(int j = 0; j < n; ++j) d[j] = -J[b + static_cast<size_t>(j) * n];
      break;
    }
    default:
      assert(false);
  }

  // z = J(:, q:) d(q:) : z[i] = dot4_{j=q..n-1}( J(i,j), d[j] ), accumulator index (j-q)&3.
  double * a0 = acc_.data();
  double * a1 = a0 + n;
  double * a2 = a1 + n;
  double * a3 = a2 + n;
  for(int i = 0; i < 4 * n; ++i) a0[i] = 0;
  double * accs[4] = {a0, a1, a2, a3};
  for(int j = q; j < n; ++j)
  {
    double * acc = accs[(j - q) & 3];
    const double * Jj = J + static_cast<size_t>(j) * n;
    double dj = d[j];
    for(int i = 0; i < n; ++i) acc[i] = std::fma(Jj[i], dj, acc[i]);
  }
  for(int i = 0; i < n; ++i) z[i] = (a0[i] + a1[i]) + (a2[i] + a3[i]);

  // r = R^-1 d(0:q) (upper triangular, column-oriented back substitution, true division)
  for(int k = 0; k < q; ++k) w[k] = d[k];
  for(int k = q - 1; k >= 0; --k)
  {
    double rk = w[k] / R[k + static_cast<size_t>(k) * n];
    r[k] = rk;
    const double * Rk = R + static_cast<size_t>(k) * n;
    for(int j = 0; j < k; ++j) w[j] = std::fma(-rk, Rk[j], w[j]);
  }

  if(instrument_)
  {
    double dn = n;
    flops_ += (sc.st < LOWER_BOUND ? 2.0 * dn * dn : 0.0) + 2.0 * dn * (dn - q) + double(q) * q;
  }
}

double GIOracle::normalDot(Selected sc, const double * v) const
{
  // ConstraintNormal::dot (include/jrl-qp/internal/ConstraintNormal.h:105-123)
  switch(sc.st)
  {
    case EQUALITY:
    case LOWER:
      return dot4(n_, C_ + static_cast<size_t>(sc.p) * ldc_, 1, v, 1);
    case UPPER:
      return -dot4(n_, C_ + static_cast<size_t>(sc.p) * ldc_, 1, v, 1);
    case FIXED:
    case LOWER_BOUND:
      return v[sc.p - A_.nbCstr()];
    case UPPER_BOUND:
      return -v[sc.p - A_.nbCstr()];
    default:
      assert(false);
      return 0;
  }
}

void GIOracle::computeStepLength(Selected sc, double & t1, double & t2, int & l)
{
  // src/GoldfarbIdnaniSolver.cpp:150-219
  const int n = n_;
  const int q = A_.nbActiveCstr();
  const double * x = x_.data();
  const double * z = z_.data();
  const double * u = u_.data();
  const double * r = r_.data();
  t1 = opt_.bigBnd;
  t2 = opt_.bigBnd;
  l = 0;
  for(int k = 0; k < q; ++k)
  {
    // NOTE: reproduces the reference's indexing quirk: the status vector is indexed by
    // the POSITION k in the active list, not by the constraint index A_[k]
    // (src/GoldfarbIdnaniSolver.cpp:162 vs src/internal/ActiveSet.cpp:71-75).
    ActivationStatus sk = A_.activationStatus(k);
    if(sk != EQUALITY && sk != FIXED)
    {
      if(instrument_) noteMargin(r[k], 0.0);
      if(r[k] > 0)
      {
        double tk = u[k] / r[k];
        if(instrument_) noteMargin(tk, t1);
        if(tk < t1)
        {
          t1 = tk;
          l = k;
        }
      }
    }
  }

  double znorm = std::sqrt(dot32(n, z, z));
  if(instrument_) noteMargin(znorm, 1e-14);
  if(znorm > 1e-14)
  {
    double b, cx, cz;
    switch(sc.st)
    {
      case LOWER:
      case UPPER:
      {
        const double * c = C_ + static_cast<size_t>(sc.p) * ldc_;
        b = sc.st == LOWER ? bl_[sc.p] : bu_[sc.p];
        cx = dot4(n, c, 1, x, 1);
        cz = dot4(n, c, 1, z, 1);
        break;
      }
      case LOWER_BOUND:
      case UPPER_BOUND:
      {
        int pb = sc.p - A_.nbCstr();
        b = sc.st == LOWER_BOUND ? xl_[pb] : xu_[pb];
        cx = x[pb];
        cz = z[pb];
        break;
      }
      default:
        assert(false);
        b = cx = cz = 0;
    }
    t2 = (b - cx) / cz;
  }
  if(instrument_) flops_ += (sc.st < LOWER_BOUND ? 4.0 * n : 0.0) + 2.0 * n + q;
}

void GIOracle::addConstraint(Selected sc)
{
  // src/DualSolver.cpp:231-235 + src/GoldfarbIdnaniSolver.cpp:221-237
  A_.activate(sc.p, sc.st);
  const int n = n_;
  const int q = A_.nbActiveCstr(); // counts the new constraint
  double * d = d_.data();
  double * J = J_.data();
  for(int i = n - 2; i >= q - 1; --i)
  {
    double c, s, rr;
    makeGivens(d[i], d[i + 1], c, s, rr);
    d[i] = rr;
    rotate(n, J + static_cast<size_t>(i) * n, 1, J + static_cast<size_t>(i + 1) * n, 1, c, s);
  }
  double * R = R_.data();
  for(int k = 0; k < q; ++k) R[k + static_cast<size_t>(q - 1) * n] = d[k];
  if(instrument_) flops_ += 6.0 * n * (n - q);
}

void GIOracle::removeConstraint(int l)
{
  // src/DualSolver.cpp:237-244
  int q = A_.nbActiveCstr();
  double * u = u_.data();
  for(int k = l; k < q; ++k) u[k] = u[k + 1];
  A_.deactivate(l);
  removeConstraintCore(l);
}

void GIOracle::removeConstraintCore(int l)
{
  // src/GoldfarbIdnaniSolver.cpp:239-256 (removeConstraint_)
  const int n = n_;
  int q = A_.nbActiveCstr(); // after removal
  double * J = J_.data();
  double * R = R_.data();
  for(int i = l; i < q; ++i)
  {
    double * Ri = R + static_cast<size_t>(i) * n;
    double * Ri1 = R + static_cast<size_t>(i + 1) * n;
    for(int k = 0; k < i; ++k) Ri[k] = Ri1[k];
    double c, s, rr;
    makeGivens(Ri1[i], Ri1[i + 1], c, s, rr);
    Ri[i] = rr;
    // rows i, i+1 of columns i+2..q
    if(q - i - 1 > 0) rotate(q - i - 1, R + i + static_cast<size_t>(i + 2) * n, n, R + (i + 1) + static_cast<size_t>(i + 2) * n, n, c, s);
    rotate(n, J + static_cast<size_t>(i) * n, 1, J + static_cast<size_t>(i + 1) * n, 1, c, s);
    if(instrument_) flops_ += 6.0 * n + 6.0 * (q - i - 1);
  }
}

const double * GIOracle::multipliers()
{
  // src/DualSolver.cpp:38-69
  int m = A_.nbAll();
  if(needExpand_)
  {
    needExpand_ = false;
    int q = A_.nbActiveCstr();
    for(int i = 0; i < m; ++i) uExp_[static_cast<size_t>(i)] = 0;
    for(int k = 0; k < q; ++k)
    {
      int i = A_[k];
      ActivationStatus s = A_.activationStatus(i);
      uExp_[static_cast<size_t>(i)] = (s == UPPER || s == UPPER_BOUND) ? u_[static_cast<size_t>(k)] : -u_[static_cast<size_t>(k)];
    }
  }
  return uExp_.data();
}

} // namespace gi_oracle
