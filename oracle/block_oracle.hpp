// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// CPU oracle for the structured dual active-set solver of jrl-umi3218/jrl-qp,
// experimental::BlockGISolver: Goldfarb-Idnani with an IMPLICIT J = L^-T Q, where L is the
// structured Cholesky factor of G (tri-block-diagonal / block-arrow) and Q a sequence of
// Householder reflectors (one per activated constraint) and Givens sequences (one per dropped
// constraint). It is the checker of jrl-qp_b200/csrc/blockgi.cuh; only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
//
// Reference files restated here (paths relative to the reference checkout):
//   src/experimental/BlockGISolver.cpp:18-60     solve()                      -> BlockGIOracle::solve
//   src/experimental/BlockGISolver.cpp:62-109    init_ (drop loop is vacuous: the solver starts from q = 0)
//   src/experimental/BlockGISolver.cpp:111-164   selectViolatedConstraint_    -> select
//   src/experimental/BlockGISolver.cpp:166-174   computeStep_                 -> computeStep
//   src/experimental/BlockGISolver.cpp:176-243   computeStepLength_ (same activationStatus(k) quirk)
//   src/experimental/BlockGISolver.cpp:245-254   addConstraint_/removeConstraint_
//   src/experimental/BlockGISolver.cpp:256-275   dot_
//   src/experimental/BlockGISolver.cpp:293-377   processInitialActiveSet
//   src/experimental/BlockGISolver.cpp:379-452   initializeComputationData ("temp" body: reset J and QR)
//   src/experimental/BlockGISolver.cpp:454-484   initializePrimalDualPoints ("temp" body: x = -G^-1 a)
//   src/structured/StructuredQR.cpp:66-103       RSolve / add (makeHouseholder) / remove (Givens)
//   src/structured/StructuredJ.cpp:33-57         premultByJ2 / premultByJt
//   src/structured/StructuredC.cpp:57-77         col() / transposeMult
//   src/internal/OrthonormalSequence.cpp:50-124,178-196  apply(Transpose)ToTheLeft, element and sequence
//   src/DualSolver.cpp:91-168,231-244            the shared loop, add/removeConstraint (as gi_oracle.cpp)
//   StructuredG (lltInPlace, solveL with the [0;v;0] hints, solveInPlaceLTranspose): decomp_oracle.cpp
//
// The reference's initializePrimalDualPoints asserts that no constraint is active at the start
// (`assert(A_.nbActiveCstr() == 0)`, :474): the solver handles inequality-only cold starts. A problem
// whose data contain an equality (bl == bu, xl == xu) — on which a release build of the reference
// silently computes with an inconsistent state — is reported here, and by the CUDA path, as
// INCONSISTENT_INPUT. Warm start data are accepted by the reference's signature but lead to the same
// assert; they are not carried by this restatement.
//
// Canonical arithmetic (same family as gi_oracle.hpp; Eigen itself is absent, see there):
//   structured factorisation / solves   decomp_oracle.hpp
//   cx_j = C_b(:,j) . x_b               dot4 over the rows of the block that holds constraint j
//   Householder H = I - tau e e^T, e = [1; essential], applied to a segment w (both directions):
//                                       s = dot128(e, w) (128 accumulators k mod 128, ascending k; folded (c + c+32) + (c+64 + c+96)
//                                       to 32, then the dot32 butterfly); w_i = fma(-(tau s), e_i, w_i)
//   Givens sequence, Q^T direction      i ascending:  (x, y) = (w_i, w_i+1): x' = fma(c,x,-(s y)); y' = fma(c,y,s x)
//   Givens sequence, Q direction        i descending: x' = fma(c,x,s y); y' = fma(c,y,-(s x))
//   makeHouseholder(d_tail)             tailSq = dot32(tail, tail); if tailSq <= DBL_MIN: tau = 0, beta = c0,
//                                       essential = 0; else beta = -sign(c0) sqrt(fma(c0,c0,tailSq)),
//                                       essential_i = tail_i / (c0 - beta), tau = (beta - c0) / beta
//   r = R^-1 d1, R updates on a drop, step length, x / u / f updates: as gi_oracle.cpp
// PARITY PINNING: tests/test_block_oracle.py checks this restatement the way the reference's own
// tests do (tests/BlockGISolverTest.in.cpp:68-123,125-170,172-230,273-310): same termination status
// and solution within 1e-8 of the dense solver on random tri-block-diagonal and arrow problems and on
// the two MultiIK fixtures (tests/golden/).
#pragma once

#include "decomp_oracle.hpp"
#include "gi_oracle.hpp"

#include <vector>

namespace block_oracle
{

using gi_oracle::ActivationStatus;
using gi_oracle::SolverOptions;
using gi_oracle::TerminationStatus;

// structured::StructuredC (block-diagonal): block i is nvar[i] x ncstr[i], column-major.
struct CBlock
{
  const double * p = nullptr;
  int rows = 0, cols = 0, ld = 0;
};

class BlockGIOracle
{
public:
  /** G: blocks of one instance (views into caller memory, factorised IN PLACE as in the
   * reference); C: block-diagonal constraint matrix; xl/xu == nullptr <=> no bounds. */
  TerminationStatus solve(decomp_oracle::Type type,
                          const std::vector<decomp_oracle::Block> & diag,
                          const std::vector<decomp_oracle::Block> & offDiag,
                          const double * a,
                          const std::vector<CBlock> & C,
                          const double * bl,
                          const double * bu,
                          const double * xl,
                          const double * xu);
  void options(const SolverOptions & o) { opt_ = o; }

  const double * solution() const { return x_.data(); }
  const double * multipliers(); // expanded, signed (src/DualSolver.cpp:38-69)
  double objectiveValue() const { return f_; }
  int iterations() const { return it_; }
  const std::vector<ActivationStatus> & activeSet() const { return A_.activationStatus(); }
  const std::vector<int> & activeList() const { return A_.activeList(); }
  int nbVar() const { return n_; }
  int nbCstr() const { return mc_; }
  long qDoubles() const { return static_cast<long>(qdata_.size()); } // size of the stored sequence
  // ---- test hooks: the OrthonormalSequence alone (tests/InternalTest.cpp:35-323 restated in tests/test_orthonormal_sequence.py)
  void seqReset(int n);
  void seqAddHouseholder(int start, int len, const double * essential, double tau); // H = I - tau e e^T, e = [1; essential(len-1)]
  void seqAddGivens(int start, int count, const double * c, const double * s); // rotations (start+i, start+i+1), i = 0 .. count-1
  void seqApply(double * v, bool transpose) const { transpose ? applyQt(v) : applyQ(v); }
  // makeHouseholder as StructuredQR::add does it on d.tail(len) (src/structured/StructuredQR.cpp:75): returns tau, beta, essential
  static void makeHouseholder(const double * x, int len, double * essential, double & tau, double & beta);
  int qRecords() const { return static_cast<int>(seq_.size()); }

private:
  struct Selected
  {
    int p = -1;
    ActivationStatus st = gi_oracle::INACTIVE;
  };
  struct Rec // one ElemOrthonormalSequence embedded at `start`
  {
    int type; // 0 Householder (size = length of e), 1 Givens (size = number of rotations)
    int start, size;
    size_t off; // into qdata_: Householder [tau, essential(size-1)], Givens [c(size), s(size)]
  };
  Selected select();
  void computeStep(Selected sc);
  void computeStepLength(Selected sc, double & t1, double & t2, int & l);
  double normalDot(Selected sc, const double * v) const;
  double colDot(int p, const double * v) const;
  void addConstraint(Selected sc);
  void removeConstraint(int l);
  void applyQt(double * v) const;
  void applyQ(double * v) const;

  SolverOptions opt_;
  int n_ = 0, mc_ = 0, nb_ = 0;
  gi_oracle::ActiveSet A_;
  decomp_oracle::StructuredG G_;
  std::vector<CBlock> C_;
  std::vector<int> cumVar_, cumCstr_, toBlock_;
  const double *a_ = nullptr, *bl_ = nullptr, *bu_ = nullptr, *xl_ = nullptr, *xu_ = nullptr;
  std::vector<double> x_, z_, d_, w_, u_, r_, R_, uExp_, qdata_;
  std::vector<Rec> seq_;
  int q_ = 0; // StructuredQR::q_
  double f_ = 0;
  int it_ = 0;
  bool needExpand_ = false;
};

} // namespace block_oracle
