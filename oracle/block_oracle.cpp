// TEST INFRASTRUCTURE — NOT PRODUCT CODE. See block_oracle.hpp for the scope, the reference
// files restated and the canonical arithmetic order.
#include "block_oracle.hpp"

#include <algorithm>
#include <atomic>
#include <cassert>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <thread>

namespace block_oracle
{

using namespace gi_oracle;

TerminationStatus BlockGIOracle::solve(decomp_oracle::Type type,
                                       const std::vector<decomp_oracle::Block> & diag,
                                       const std::vector<decomp_oracle::Block> & offDiag,
                                       const double * a,
                                       const std::vector<CBlock> & C,
                                       const double * bl,
                                       const double * bu,
                                       const double * xl,
                                       const double * xu)
{
  // src/experimental/BlockGISolver.cpp:18-60
  G_ = decomp_oracle::StructuredG(type, diag, offDiag);
  n_ = G_.nbVar();
  C_ = C;
  cumVar_.clear();
  cumCstr_.clear();
  toBlock_.clear();
  int nv = 0, nc = 0;
  for(size_t i = 0; i < C.size(); ++i) // StructuredC::StructuredC (src/structured/StructuredC.cpp:9-25)
  {
    cumVar_.push_back(nv);
    cumCstr_.push_back(nc);
    for(int k = 0; k < C[i].cols; ++k) toBlock_.push_back(static_cast<int>(i));
    nv += C[i].rows;
    nc += C[i].cols;
  }
  cumVar_.push_back(nv);
  cumCstr_.push_back(nc);
  assert(nv == n_ || C.empty());
  mc_ = nc;
  nb_ = xl ? n_ : 0;
  a_ = a;
  bl_ = bl;
  bu_ = bu;
  xl_ = xl;
  xu_ = xu;
  const size_t n = static_cast<size_t>(n_);
  A_.resize(mc_, nb_);
  x_.assign(n, 0);
  z_.assign(n, 0);
  d_.assign(n, 0);
  w_.assign(n, 0);
  u_.assign(n + 1, 0);
  r_.assign(n + 1, 0);
  R_.assign(n * n, 0);
  uExp_.assign(static_cast<size_t>(mc_ + nb_), 0);
  seq_.clear();
  qdata_.clear();
  q_ = 0;
  f_ = 0;
  it_ = 0;
  needExpand_ = true;

  // ---- init_ (src/experimental/BlockGISolver.cpp:62-109)
  // processInitialActiveSet (:293-377), cold start: equalities of the data are activated ...
  A_.reset();
  for(int i = 0; i < nb_; ++i)
    if(xl_[i] == xu_[i]) A_.activate(mc_ + i, FIXED);
  for(int i = 0; i < mc_; ++i)
    if(bl_[i] == bu_[i]) A_.activate(i, EQUALITY);
  if(A_.nbActiveCstr() > n_ && A_.nbActiveEquality() + A_.nbFixedVariable() > n_) return OVERCONSTRAINED_PROBLEM;
  // ... and initializePrimalDualPoints then asserts there is none (:474). See the header.
  if(A_.nbActiveCstr() > 0)
  {
    A_.reset();
    return INCONSISTENT_INPUT;
  }
  if(!G_.lltInPlace()) return NON_POS_HESSIAN; // :71-73
  // initializeComputationData: J_.reset(); QR_.reset(); J = (L, Q = I)
  // initializePrimalDualPoints (:476-481): x = -G^-1 a, f = 0.5 a.x
  double * x = x_.data();
  G_.solveL(x, a_);
  G_.solveInPlaceLTranspose(x);
  for(int i = 0; i < n_; ++i) x[i] = -x[i];
  f_ = 0.5 * dot32(n_, a_, x);

  // ---- DualSolver::solve (src/DualSolver.cpp:96-168)
  bool skipStep1 = false;
  Selected sc;
  double * z = z_.data();
  double * u = u_.data();
  double * r = r_.data();
  for(; it_ < opt_.maxIter; ++it_)
  {
    int q = A_.nbActiveCstr();
    if(!skipStep1)
    {
      sc = select();
      if(sc.st == INACTIVE) return SUCCESS;
      u[q] = 0;
    }
    computeStep(sc);
    double t1, t2;
    int l;
    computeStepLength(sc, t1, t2, l);
    double t = std::min(t1, t2);
    if(t >= opt_.bigBnd) return INFEASIBLE;
    if(t2 >= opt_.bigBnd)
    {
      for(int k = 0; k < q; ++k) u[k] = std::fma(-t, r[k], u[k]);
      u[q] += t;
      removeConstraint(l);
      skipStep1 = true;
    }
    else
    {
      for(int i = 0; i < n_; ++i) x[i] = std::fma(t, z[i], x[i]);
      f_ += (t * normalDot(sc, z)) * (0.5 * t + u[q]);
      for(int k = 0; k < q; ++k) u[k] = std::fma(-t, r[k], u[k]);
      u[q] += t;
      if(t == t2)
      {
        addConstraint(sc);
        skipStep1 = false;
      }
      else
      {
        removeConstraint(l);
        skipStep1 = true;
      }
    }
  }
  return MAX_ITER_REACHED;
}

// C.col(p).dot(v): SingleNZSegmentVector::dot = dot over the rows of the block (StructuredC.cpp:57-62)
double BlockGIOracle::colDot(int p, const double * v) const
{
  const int bi = toBlock_[static_cast<size_t>(p)];
  const CBlock & B = C_[static_cast<size_t>(bi)];
  const double * c = B.p + static_cast<std::ptrdiff_t>(p - cumCstr_[static_cast<size_t>(bi)]) * B.ld;
  return dot4(B.rows, c, 1, v + cumVar_[static_cast<size_t>(bi)], 1);
}

BlockGIOracle::Selected BlockGIOracle::select()
{
  // src/experimental/BlockGISolver.cpp:111-164; cx = C^T x block by block (StructuredC::transposeMult)
  const double * x = x_.data();
  double smin = 0;
  Selected sel;
  for(int i = 0; i < mc_; ++i)
  {
    if(!A_.isActive(i))
    {
      double cx = colDot(i, x);
      double sl = cx - bl_[i];
      if(sl < smin)
      {
        smin = sl;
        sel = {i, LOWER};
      }
      else
      {
        double su = bu_[i] - cx;
        if(su < smin)
        {
          smin = su;
          sel = {i, UPPER};
        }
      }
    }
  }
  for(int i = 0; i < nb_; ++i)
  {
    if(!A_.isActiveBnd(i))
    {
      double sl = x[i] - xl_[i];
      if(sl < smin)
      {
        smin = sl;
        sel = {mc_ + i, LOWER_BOUND};
      }
      else
      {
        double su = xu_[i] - x[i];
        if(su < smin)
        {
          smin = su;
          sel = {mc_ + i, UPPER_BOUND};
        }
      }
    }
  }
  return sel;
}

// OrthonormalSequence::applyTransposeToTheLeft(VectorRef) (src/internal/OrthonormalSequence.cpp:187-196)
void BlockGIOracle::applyQt(double * v) const
{
  for(const Rec & h : seq_)
  {
    double * w = v + h.start;
    const double * p = qdata_.data() + h.off;
    if(h.type == 0)
    {
      // size_ == 1 Householder branch (:104-108): d = E.dot(w); w -= h d E, E = [1; essential]
      const double tau = p[0];
      // E.w over 128 classes (k mod 128, ascending k), folded (c, c+32), (c+64, c+96), then the dot32 butterfly
      double a128[128];
      for(int l = 0; l < 128; ++l) a128[l] = 0;
      a128[0] = std::fma(1.0, w[0], a128[0]);
      for(int k = 1; k < h.size; ++k) a128[k & 127] = std::fma(p[k], w[k], a128[k & 127]);
      double acc[32];
      for(int l = 0; l < 32; ++l) acc[l] = (a128[l] + a128[l + 32]) + (a128[l + 64] + a128[l + 96]);
      for(int off = 16; off >= 1; off >>= 1)
      {
        double nxt[32];
        for(int l = 0; l < 32; ++l) nxt[l] = acc[l] + acc[l ^ off];
        for(int l = 0; l < 32; ++l) acc[l] = nxt[l];
      }
      const double hd = tau * acc[0];
      w[0] = std::fma(-hd, 1.0, w[0]);
      for(int k = 1; k < h.size; ++k) w[k] = std::fma(-hd, p[k], w[k]);
    }
    else
    {
      // Givens(c, s).transpose(), i ascending (:117-122)
      const double * c = p;
      const double * s = p + h.size;
      for(int i = 0; i < h.size; ++i)
      {
        double xi = w[i], yi = w[i + 1];
        w[i] = std::fma(c[i], xi, -(s[i] * yi));
        w[i + 1] = std::fma(c[i], yi, s[i] * xi);
      }
    }
  }
}

// OrthonormalSequence::applyToTheLeft(VectorRef) (src/internal/OrthonormalSequence.cpp:178-185)
void BlockGIOracle::applyQ(double * v) const
{
  for(size_t e = seq_.size(); e-- > 0;)
  {
    const Rec & h = seq_[e];
    double * w = v + h.start;
    const double * p = qdata_.data() + h.off;
    if(h.type == 0)
    {
      const double tau = p[0];
      // E.w over 128 classes (k mod 128, ascending k), folded (c, c+32), (c+64, c+96), then the dot32 butterfly
      double a128[128];
      for(int l = 0; l < 128; ++l) a128[l] = 0;
      a128[0] = std::fma(1.0, w[0], a128[0]);
      for(int k = 1; k < h.size; ++k) a128[k & 127] = std::fma(p[k], w[k], a128[k & 127]);
      double acc[32];
      for(int l = 0; l < 32; ++l) acc[l] = (a128[l] + a128[l + 32]) + (a128[l + 64] + a128[l + 96]);
      for(int off = 16; off >= 1; off >>= 1)
      {
        double nxt[32];
        for(int l = 0; l < 32; ++l) nxt[l] = acc[l] + acc[l ^ off];
        for(int l = 0; l < 32; ++l) acc[l] = nxt[l];
      }
      const double hd = tau * acc[0];
      w[0] = std::fma(-hd, 1.0, w[0]);
      for(int k = 1; k < h.size; ++k) w[k] = std::fma(-hd, p[k], w[k]);
    }
    else
    {
      // Givens(c, s), i descending (:69-75)
      const double * c = p;
      const double * s = p + h.size;
      for(int i = h.size - 1; i >= 0; --i)
      {
        double xi = w[i], yi = w[i + 1];
        w[i] = std::fma(c[i], xi, s[i] * yi);
        w[i + 1] = std::fma(c[i], yi, -(s[i] * xi));
      }
    }
  }
}

void BlockGIOracle::computeStep(Selected sc)
{
  // src/experimental/BlockGISolver.cpp:166-174
  const int n = n_;
  const int q = A_.nbActiveCstr();
  double * d = d_.data();
  double * z = z_.data();
  double * r = r_.data();
  double * w = w_.data();
  // StructuredJ::premultByJt (src/structured/StructuredJ.cpp:44-57): d = Q^T L^-1 n+
  std::fill(w, w + n, 0.0);
  if(sc.st <= EQUALITY)
  {
    const int bi = toBlock_[static_cast<size_t>(sc.p)];
    const CBlock & B = C_[static_cast<size_t>(bi)];
    const int s0 = cumVar_[static_cast<size_t>(bi)];
    const double * c = B.p + static_cast<std::ptrdiff_t>(sc.p - cumCstr_[static_cast<size_t>(bi)]) * B.ld;
    for(int i = 0; i < B.rows; ++i) w[s0 + i] = c[i];
    G_.solveL(d, w, s0, s0 + B.rows);
    if(sc.st == UPPER)
      for(int i = 0; i < n; ++i) d[i] = -d[i]; // out *= -1
  }
  else
  {
    const int b = sc.p - mc_;
    w[b] = sc.st == UPPER_BOUND ? -1.0 : 1.0;
    G_.solveL(d, w, b, b + 1);
  }
  applyQt(d);
  // StructuredJ::premultByJ2 (:33-42): z = L^-T Q [0; d2]
  for(int i = 0; i < q; ++i) z[i] = 0;
  for(int i = q; i < n; ++i) z[i] = d[i];
  applyQ(z);
  G_.solveInPlaceLTranspose(z);
  // StructuredQR::RSolve (src/structured/StructuredQR.cpp:66-71): column-oriented, true division
  const double * R = R_.data();
  for(int k = 0; k < q; ++k) w[k] = d[k];
  for(int k = q - 1; k >= 0; --k)
  {
    double rk = w[k] / R[k + static_cast<size_t>(k) * n];
    r[k] = rk;
    const double * Rk = R + static_cast<size_t>(k) * n;
    for(int j = 0; j < k; ++j) w[j] = std::fma(-rk, Rk[j], w[j]);
  }
}

double BlockGIOracle::normalDot(Selected sc, const double * v) const
{
  // src/experimental/BlockGISolver.cpp:256-275
  switch(sc.st)
  {
    case EQUALITY:
    case LOWER:
      return colDot(sc.p, v);
    case UPPER:
      return -colDot(sc.p, v);
    case FIXED:
    case LOWER_BOUND:
      return v[sc.p - mc_];
    case UPPER_BOUND:
      return -v[sc.p - mc_];
    default:
      assert(false);
      return 0;
  }
}

void BlockGIOracle::computeStepLength(Selected sc, double & t1, double & t2, int & l)
{
  // src/experimental/BlockGISolver.cpp:176-243
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
    ActivationStatus sk = A_.activationStatus(k); // the reference's indexing quirk (:188), kept
    if(sk != EQUALITY && sk != FIXED && r[k] > 0)
    {
      double tk = u[k] / r[k];
      if(tk < t1)
      {
        t1 = tk;
        l = k;
      }
    }
  }
  double znorm = std::sqrt(dot32(n_, z, z));
  if(znorm > 1e-14)
  {
    double b, cx, cz;
    switch(sc.st)
    {
      case LOWER:
      case UPPER:
        b = sc.st == LOWER ? bl_[sc.p] : bu_[sc.p];
        cx = colDot(sc.p, x);
        cz = colDot(sc.p, z);
        break;
      case LOWER_BOUND:
      case UPPER_BOUND:
      {
        int pb = sc.p - mc_;
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
}

void BlockGIOracle::addConstraint(Selected sc)
{
  // DualSolver::addConstraint (src/DualSolver.cpp:231-235) + StructuredQR::add (StructuredQR.cpp:72-86)
  A_.activate(sc.p, sc.st);
  const int n = n_;
  const int q = q_;
  const double * d = d_.data();
  const int len = n - q; // d.tail(n - q)
  const double c0 = d[q];
  const double * tail = d + q + 1;
  const double tailSq = len == 1 ? 0.0 : dot32(len - 1, tail, tail);
  Rec h{0, q, len, qdata_.size()};
  qdata_.resize(qdata_.size() + static_cast<size_t>(len));
  double * p = qdata_.data() + h.off;
  double tau, beta;
  if(tailSq <= DBL_MIN)
  {
    tau = 0;
    beta = c0;
    for(int i = 1; i < len; ++i) p[i] = 0;
  }
  else
  {
    beta = std::sqrt(std::fma(c0, c0, tailSq));
    if(c0 >= 0) beta = -beta;
    const double den = c0 - beta;
    for(int i = 1; i < len; ++i) p[i] = tail[i - 1] / den;
    tau = (beta - c0) / beta;
  }
  p[0] = tau;
  double * R = R_.data();
  for(int k = 0; k < q; ++k) R[k + static_cast<size_t>(q) * n] = d[k];
  R[q + static_cast<size_t>(q) * n] = beta;
  seq_.push_back(h);
  ++q_;
}

void BlockGIOracle::removeConstraint(int l)
{
  // DualSolver::removeConstraint (src/DualSolver.cpp:237-244)
  int qa = A_.nbActiveCstr();
  double * u = u_.data();
  for(int k = l; k < qa; ++k) u[k] = u[k + 1];
  A_.deactivate(l);
  // StructuredQR::remove (src/structured/StructuredQR.cpp:88-103)
  --q_;
  const int n = n_;
  const int q = q_;
  double * R = R_.data();
  const int g = q - l;
  if(g <= 0) return; // an empty Givens sequence: nothing is ever applied
  Rec h{1, l, g, qdata_.size()};
  qdata_.resize(qdata_.size() + 2 * static_cast<size_t>(g));
  double * cs = qdata_.data() + h.off;
  for(int i = l; i < q; ++i)
  {
    double * Ri = R + static_cast<size_t>(i) * n;
    double * Ri1 = R + static_cast<size_t>(i + 1) * n;
    for(int k = 0; k < i; ++k) Ri[k] = Ri1[k];
    double c, s, rr;
    makeGivens(Ri1[i], Ri1[i + 1], c, s, rr);
    Ri[i] = rr;
    for(int j = i + 2; j <= q; ++j) // R.rightCols(q - i - 1).applyOnTheLeft(i, i+1, Qi^T)
    {
      double * Rj = R + static_cast<size_t>(j) * n;
      double xi = Rj[i], yi = Rj[i + 1];
      Rj[i] = std::fma(c, xi, -(s * yi));
      Rj[i + 1] = std::fma(c, yi, s * xi);
    }
    cs[i - l] = c;
    cs[g + i - l] = s;
  }
  seq_.push_back(h);
}

const double * BlockGIOracle::multipliers()
{
  // src/DualSolver.cpp:38-69
  if(needExpand_)
  {
    needExpand_ = false;
    std::fill(uExp_.begin(), uExp_.end(), 0.0);
    int q = A_.nbActiveCstr();
    for(int k = 0; k < q; ++k)
    {
      int i = A_[k];
      ActivationStatus s = A_.activationStatus(i);
      uExp_[static_cast<size_t>(i)] = (s == UPPER || s == UPPER_BOUND) ? u_[static_cast<size_t>(k)] : -u_[static_cast<size_t>(k)];
    }
  }
  return uExp_.data();
}


// ---- test hooks (see block_oracle.hpp)
void BlockGIOracle::seqReset(int n)
{
  n_ = n;
  seq_.clear();
  qdata_.clear();
}

void BlockGIOracle::seqAddHouseholder(int start, int len, const double * essential, double tau)
{
  Rec h{0, start, len, qdata_.size()};
  qdata_.resize(qdata_.size() + static_cast<size_t>(len));
  double * p = qdata_.data() + h.off;
  p[0] = tau;
  for(int i = 1; i < len; ++i) p[i] = essential[i - 1];
  seq_.push_back(h);
}

void BlockGIOracle::seqAddGivens(int start, int count, const double * c, const double * s)
{
  Rec h{1, start, count, qdata_.size()};
  qdata_.resize(qdata_.size() + 2 * static_cast<size_t>(count));
  double * p = qdata_.data() + h.off;
  for(int i = 0; i < count; ++i)
  {
    p[i] = c[i];
    p[count + i] = s[i];
  }
  seq_.push_back(h);
}

void BlockGIOracle::makeHouseholder(const double * x, int len, double * essential, double & tau, double & beta)
{
  const double c0 = x[0];
  const double tailSq = len == 1 ? 0.0 : dot32(len - 1, x + 1, x + 1);
  if(tailSq <= DBL_MIN)
  {
    tau = 0;
    beta = c0;
    for(int i = 1; i < len; ++i) essential[i - 1] = 0;
  }
  else
  {
    beta = std::sqrt(std::fma(c0, c0, tailSq));
    if(c0 >= 0) beta = -beta;
    const double den = c0 - beta;
    for(int i = 1; i < len; ++i) essential[i - 1] = x[i] / den;
    tau = (beta - c0) / beta;
  }
}

} // namespace block_oracle

// ---------------------------------------------------------------------------------------------
// C entry point (ctypes). Structure descriptors as include/jrlqp_b200.h: jrlqp_structure for G,
// jrlqp_cstructure for the block-diagonal C.
// ---------------------------------------------------------------------------------------------
extern "C" int block_oracle_solve_batch(int type,
                                        int b,
                                        const int * size,
                                        const long * doff,
                                        const int * dld,
                                        const long * ooff,
                                        const int * old,
                                        long gspan, // elements of one instance of G (copied, factorised privately)
                                        int cb,
                                        const int * cnvar,
                                        const int * cncstr,
                                        const long * coff,
                                        const int * cld,
                                        int use_bounds,
                                        long batch,
                                        const double * G,
                                        long sG,
                                        const double * a,
                                        long sa,
                                        const double * C,
                                        long sC,
                                        const double * bl,
                                        long sbl,
                                        const double * bu,
                                        long sbu,
                                        const double * xl,
                                        long sxl,
                                        const double * xu,
                                        long sxu,
                                        int max_iter,
                                        double big_bnd,
                                        double * x,
                                        double * u,
                                        double * f,
                                        int * iters,
                                        int * status,
                                        signed char * act,
                                        int * active_list,
                                        int * nactive,
                                        double * L_out, // nullable: [batch][gspan], the factorised blocks
                                        long * qdoubles, // nullable: size of the stored orthonormal sequence
                                        int nthreads)
{
  using namespace block_oracle;
  int n = 0, mc = 0;
  for(int i = 0; i < b; ++i) n += size[i];
  for(int i = 0; i < cb; ++i) mc += cncstr[i];
  const int nb = use_bounds ? n : 0;
  const int m = mc + nb;
  if(nthreads < 1) nthreads = 1;
  if(nthreads > batch) nthreads = static_cast<int>(std::max<long>(1, batch));
  std::atomic<long> next(0);
  std::atomic<int> worst(0);
  auto worker = [&]()
  {
    BlockGIOracle solver;
    gi_oracle::SolverOptions opt;
    opt.maxIter = max_iter;
    opt.bigBnd = big_bnd;
    solver.options(opt);
    std::vector<double> Gs(static_cast<size_t>(gspan));
    std::vector<decomp_oracle::Block> diag(static_cast<size_t>(b)), off(static_cast<size_t>(std::max(0, b - 1)));
    std::vector<CBlock> Cb(static_cast<size_t>(cb));
    int localWorst = 0;
    for(;;)
    {
      long k0 = next.fetch_add(4);
      if(k0 >= batch) break;
      for(long k = k0; k < std::min(batch, k0 + 4); ++k)
      {
        std::memcpy(Gs.data(), G + k * sG, sizeof(double) * static_cast<size_t>(gspan));
        for(int i = 0; i < b; ++i) diag[static_cast<size_t>(i)] = {Gs.data() + doff[i], size[i], size[i], dld[i]};
        for(int i = 0; i + 1 < b; ++i)
        {
          int rows, cols;
          if(type == decomp_oracle::TriBlockDiagonal)
          {
            rows = size[i + 1];
            cols = size[i];
          }
          else if(type == decomp_oracle::BlockArrowDown)
          {
            rows = size[b - 1];
            cols = size[i];
          }
          else
          {
            rows = size[i + 1];
            cols = size[0];
          }
          off[static_cast<size_t>(i)] = {Gs.data() + ooff[i], rows, cols, old[i]};
        }
        for(int i = 0; i < cb; ++i) Cb[static_cast<size_t>(i)] = {C + k * sC + coff[i], cnvar[i], cncstr[i], cld[i]};
        int st = solver.solve(static_cast<decomp_oracle::Type>(type), diag, off, a + k * sa, Cb, bl + k * sbl, bu + k * sbu,
                              nb ? xl + k * sxl : nullptr, nb ? xu + k * sxu : nullptr);
        localWorst = std::max(localWorst, st);
        const bool failed = st == NON_POS_HESSIAN || st == OVERCONSTRAINED_PROBLEM || st == INCONSISTENT_INPUT;
        if(x)
        {
          if(failed)
            std::memset(x + k * n, 0, sizeof(double) * static_cast<size_t>(n));
          else
            std::memcpy(x + k * n, solver.solution(), sizeof(double) * static_cast<size_t>(n));
        }
        if(u)
        {
          if(failed)
            std::memset(u + k * m, 0, sizeof(double) * static_cast<size_t>(m));
          else
            std::memcpy(u + k * m, solver.multipliers(), sizeof(double) * static_cast<size_t>(m));
        }
        if(f) f[k] = failed ? 0 : solver.objectiveValue();
        if(iters) iters[k] = failed ? 0 : solver.iterations();
        if(status) status[k] = st;
        const auto & as = solver.activeSet();
        if(act)
          for(int i = 0; i < m; ++i) act[k * m + i] = failed ? 0 : static_cast<signed char>(as[static_cast<size_t>(i)]);
        const auto & al = solver.activeList();
        if(active_list)
          for(int i = 0; i < n; ++i) active_list[k * n + i] = (!failed && i < static_cast<int>(al.size())) ? al[static_cast<size_t>(i)] : -1;
        if(nactive) nactive[k] = failed ? 0 : static_cast<int>(al.size());
        if(L_out) std::memcpy(L_out + k * gspan, Gs.data(), sizeof(double) * static_cast<size_t>(gspan));
        if(qdoubles) qdoubles[k] = solver.qDoubles();
      }
    }
    int w = worst.load();
    while(localWorst > w && !worst.compare_exchange_weak(w, localWorst)) {}
  };
  if(nthreads == 1)
    worker();
  else
  {
    std::vector<std::thread> th;
    for(int t = 0; t < nthreads; ++t) th.emplace_back(worker);
    for(auto & t : th) t.join();
  }
  return worst.load();
}

// ---- test hooks: OrthonormalSequence alone
extern "C" void * block_oracle_seq_create(int n)
{
  auto * o = new block_oracle::BlockGIOracle();
  o->seqReset(n);
  return o;
}
extern "C" void block_oracle_seq_destroy(void * h)
{
  delete static_cast<block_oracle::BlockGIOracle *>(h);
}
extern "C" void block_oracle_seq_add_householder(void * h, int start, int len, const double * essential, double tau)
{
  static_cast<block_oracle::BlockGIOracle *>(h)->seqAddHouseholder(start, len, essential, tau);
}
extern "C" void block_oracle_seq_add_givens(void * h, int start, int count, const double * c, const double * s)
{
  static_cast<block_oracle::BlockGIOracle *>(h)->seqAddGivens(start, count, c, s);
}
extern "C" void block_oracle_seq_apply(void * h, double * v, int transpose)
{
  static_cast<block_oracle::BlockGIOracle *>(h)->seqApply(v, transpose != 0);
}
extern "C" void block_oracle_make_householder(const double * x, int len, double * essential, double * tau, double * beta)
{
  block_oracle::BlockGIOracle::makeHouseholder(x, len, essential, *tau, *beta);
}
