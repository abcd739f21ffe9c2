// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// CPU oracle: a plain C++ restatement of the reference's dense Goldfarb-Idnani
// dual active-set solver (jrl-umi3218/jrl-qp), used only as the checker for the
// CUDA path (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline /
// --impl reference legs). Nothing under jrl-qp_b200/ may include or link this.
//
// Reference files restated here (paths relative to the reference checkout):
//   src/DualSolver.cpp:38-69      multipliers()  -> GIOracle::expandMultipliers
//   src/DualSolver.cpp:91-168     solve() loop   -> GIOracle::solve
//   src/DualSolver.cpp:231-244    add/removeConstraint
//   src/GoldfarbIdnaniSolver.cpp:56-82    init_            -> GIOracle::init
//   src/GoldfarbIdnaniSolver.cpp:84-134   selectViolated   -> GIOracle::select
//   src/GoldfarbIdnaniSolver.cpp:136-148  computeStep_     -> GIOracle::computeStep
//   src/GoldfarbIdnaniSolver.cpp:150-219  computeStepLength_ (incl. the
//                                         activationStatus(k) indexing quirk)
//   src/GoldfarbIdnaniSolver.cpp:221-237  addConstraint_   (Givens sweep)
//   src/GoldfarbIdnaniSolver.cpp:239-256  removeConstraint_
//   src/GoldfarbIdnaniSolver.cpp:268-338  initActiveSet / addInitialConstraint
//   src/internal/ActiveSet.cpp:45-168     ActiveSet
//   include/jrl-qp/internal/ConstraintNormal.h:81-123  preMultiplyByMt / dot
//
// Third-party arithmetic: the reference delegates its linear algebra to Eigen 3
// (un-vendored, un-pinned: README.md:37-38 ">= 3.2.8"), which is absent from
// /root/reference and from this image, so the reference itself cannot be
// compiled here. The Eigen primitives on the path (llt_inplace, triangular
// solves, gemv/dot, JacobiRotation::makeGivens, applyOnTheRight/Left) are
// restated from their published algorithms. Eigen's floating-point summation
// order depends on its version, SIMD width and FMA contraction, none of which
// the reference pins; this oracle therefore fixes ONE canonical operation
// order (documented per function, and in DESIGN.md) that the CUDA kernels
// reproduce bit for bit. PARITY PINNING: the oracle is checked against every
// known-answer test the reference holds for this path (tests/golden/, see
// tests/test_oracle_golden.py); bit-level parity with an Eigen build of jrl-qp
// is not pinnable here and is documented as such.
//
// Canonical arithmetic (all IEEE-754 binary64, round-to-nearest-even, no
// implicit contraction — compile with -ffp-contract=off):
//   dot4(len, a, b)   4 interleaved accumulators acc[k&3] = fma(a[k],b[k],acc[k&3]),
//                     k ascending; result (acc0+acc1)+(acc2+acc3).
//   dot32(len, a, b)  32 interleaved accumulators acc[k&31], then the xor
//                     butterfly acc[l] += acc[l^off], off = 16,8,4,2,1.
//   axpy updates      y = fma(alpha, x, y).
//   Givens            Eigen's JacobiRotation::makeGivens (real case) verbatim in
//                     exact arithmetic order; apply: x' = fma(c,x,-(s*y)),
//                     y' = fma(c,y,s*x).
#pragma once

#include <cstddef>
#include <cstdint>
#include <vector>

namespace gi_oracle
{

// include/jrl-qp/enums.h:14-23 (order matters)
enum ActivationStatus : int8_t
{
  INACTIVE = 0,
  LOWER = 1,
  UPPER = 2,
  EQUALITY = 3,
  LOWER_BOUND = 4,
  UPPER_BOUND = 5,
  FIXED = 6
};

// include/jrl-qp/enums.h:26-37
enum TerminationStatus : int
{
  SUCCESS = 0,
  INCONSISTENT_INPUT = 1,
  NON_POS_HESSIAN = 2,
  INFEASIBLE = 3,
  MAX_ITER_REACHED = 4,
  LINEAR_DEPENDENCY_DETECTED = 5,
  OVERCONSTRAINED_PROBLEM = 6,
  UNKNOWN = 7
};

// include/jrl-qp/SolverOptions.h:14-22 (logging fields carried for API parity only)
struct SolverOptions
{
  int maxIter = 500;
  double bigBnd = 1e100;
  bool warmStart = false;
  std::uint32_t logFlags = 0;
};

double dot4(int len, const double * a, std::ptrdiff_t sa, const double * b, std::ptrdiff_t sb);
double dot32(int len, const double * a, const double * b);
// Eigen JacobiRotation::makeGivens(p, q, &r), real case. G = [c s; -s c], G^T [p;q] = [r;0].
void makeGivens(double p, double q, double & c, double & s, double & r);

// src/internal/ActiveSet.cpp restated.
class ActiveSet
{
public:
  ActiveSet(int nCstr = 0, int nBnd = 0) { resize(nCstr, nBnd); }
  void resize(int nCstr, int nBnd);
  void reset();
  bool isActive(int i) const { return status_[static_cast<size_t>(i)] != INACTIVE; }
  bool isActiveBnd(int i) const { return status_[static_cast<size_t>(nbCstr_ + i)] != INACTIVE; }
  ActivationStatus activationStatus(int i) const { return status_[static_cast<size_t>(i)]; }
  const std::vector<ActivationStatus> & activationStatus() const { return status_; }
  int operator[](int k) const { return activeSet_[static_cast<size_t>(k)]; }
  const std::vector<int> & activeList() const { return activeSet_; }
  void activate(int cstrIdx, ActivationStatus status);
  void deactivate(int activeIdx);
  int nbCstr() const { return nbCstr_; }
  int nbBnd() const { return nbBnd_; }
  int nbAll() const { return nbCstr_ + nbBnd_; }
  int nbActiveCstr() const { return me_ + mi_ + mb_; }
  int nbActiveEquality() const { return me_; }
  int nbActiveInequality() const { return mi_; }
  int nbActiveLowerInequality() const { return ml_; }
  int nbActiveUpperInequality() const { return mu_; }
  int nbActiveBound() const { return mb_; }
  int nbActiveLowerBound() const { return mbl_; }
  int nbActiveUpperBound() const { return mbu_; }
  int nbFixedVariable() const { return mbe_; }

private:
  void count(ActivationStatus s, int delta);
  std::vector<ActivationStatus> status_;
  std::vector<int> activeSet_;
  int nbCstr_ = 0, nbBnd_ = 0;
  int me_ = 0, mi_ = 0, ml_ = 0, mu_ = 0, mb_ = 0, mbl_ = 0, mbu_ = 0, mbe_ = 0;
};

struct TraceEvent
{
  int it; // iteration index (-1 for equality pre-activation)
  int p; // selected constraint index
  int status; // ActivationStatus of the selected constraint
  int l; // drop candidate (position in active list)
  int kind; // 0 = add (full step), 1 = drop after partial step, 2 = drop after dual-only step, 3 = pre-activation
  double t1, t2;
};

/** Dense cold-start Goldfarb-Idnani solver, one QP per call, mirroring
 * jrl::qp::GoldfarbIdnaniSolver (include/jrl-qp/GoldfarbIdnaniSolver.h:15-33).
 * G: n x n column-major, leading dimension ldg, lower triangle read and
 * overwritten by its Cholesky factor. C: n x mc column-major (one constraint
 * per column), leading dimension ldc. xl/xu == nullptr <=> no bounds.
 */
class GIOracle
{
public:
  GIOracle(int n = 0, int mc = 0, bool useBounds = false);
  void resize(int n, int mc, bool useBounds);
  void options(const SolverOptions & o) { opt_ = o; }
  const SolverOptions & options() const { return opt_; }
  void instrument(bool on) { instrument_ = on; }

  TerminationStatus solve(double * G,
                          int ldg,
                          const double * a,
                          const double * C,
                          int ldc,
                          const double * bl,
                          const double * bu,
                          const double * xl,
                          const double * xu);

  /** experimental::GoldfarbIdnaniSolver::solve (src/experimental/GoldfarbIdnaniSolver.cpp:21-64), the
   * warm-start capable variant: `as` (nullable, nbCstr + nbBnd entries, general constraints first) is the
   * initial guess of the active set; with options().warmStart and as == nullptr the previous active set
   * of this object is reused. See warm_oracle.cpp for the restated functions and arithmetic. */
  TerminationStatus solveExperimental(double * G,
                                      int ldg,
                                      const double * a,
                                      const double * C,
                                      int ldc,
                                      const double * bl,
                                      const double * bu,
                                      const double * xl,
                                      const double * xu,
                                      const int8_t * as);

  const double * solution() const { return x_.data(); }
  const double * multipliers(); // expanded, signed (src/DualSolver.cpp:38-69)
  double objectiveValue() const { return f_; }
  int iterations() const { return it_; }
  const std::vector<ActivationStatus> & activeSet() const { return A_.activationStatus(); }
  const std::vector<int> & activeList() const { return A_.activeList(); }
  void resetActiveSet() { A_.reset(); }
  int nbVar() const { return n_; }
  int nbCstr() const { return A_.nbCstr(); }
  int nbBnd() const { return A_.nbBnd(); }

  // instrumentation (valid when instrument(true))
  double flops() const { return flops_; }
  double minMargin() const { return margin_; }
  const std::vector<TraceEvent> & trace() const { return trace_; }
  const double * J() const { return J_.data(); } // n x n col-major, ld n
  const double * R() const { return R_.data(); } // n x n col-major, ld n (upper q x q used)

private:
  struct Selected
  {
    int p = -1;
    ActivationStatus st = INACTIVE;
  };
  bool init();
  bool overconstrained_ = false; // more than nbVar equalities / fixed variables met by initActiveSet
  void initActiveSet();
  void addInitialConstraint(Selected sc);
  Selected select();
  void computeStep(Selected sc);
  void computeStepLength(Selected sc, double & t1, double & t2, int & l);
  double normalDot(Selected sc, const double * v) const; // ConstraintNormal::dot
  void addConstraint(Selected sc);
  void removeConstraint(int l);
  void removeConstraintCore(int l);
  TerminationStatus mainLoop();
  // experimental solver (warm_oracle.cpp)
  TerminationStatus initExperimental();
  TerminationStatus processInitialActiveSet();
  void initializeComputationData();
  void initializePrimalDualPoints();
  void noteMargin(double a, double b);

  SolverOptions opt_;
  int n_ = 0;
  ActiveSet A_;
  // problem views
  double * G_ = nullptr;
  int ldg_ = 0, ldc_ = 0;
  const double *a_ = nullptr, *C_ = nullptr, *bl_ = nullptr, *bu_ = nullptr, *xl_ = nullptr, *xu_ = nullptr;
  // workspaces
  std::vector<double> x_, z_, d_, w_, u_, r_, J_, R_, uExp_, acc_;
  std::vector<double> bact_, hcoef_, alpha_; // experimental solver
  std::vector<ActivationStatus> asIn_; // pb_.as
  double f_ = 0;
  int it_ = 0;
  bool needExpand_ = false;
  // instrumentation
  bool instrument_ = false;
  double flops_ = 0;
  double margin_ = 0;
  std::vector<TraceEvent> trace_;
};

} // namespace gi_oracle
