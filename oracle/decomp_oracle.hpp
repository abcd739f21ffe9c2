// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// CPU oracle for the structured Cholesky decompositions of jrl-umi3218/jrl-qp and their
// triangular solves (the checker of jrl-qp_b200/csrc/structured.cu). Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
//
// Reference files restated here (paths relative to the reference checkout):
//   src/decomposition/triBlockDiagLLT.cpp:9-36     triBlockDiagLLT
//   src/decomposition/triBlockDiagLLT.cpp:38-98    triBlockDiagLSolve (with the `start` hint)
//   src/decomposition/triBlockDiagLLT.cpp:100-158  triBlockDiagLTransposeSolve (`end` hint)
//   src/decomposition/blockArrowLLT.cpp:52-90      blockArrowLLT_<Up> / blockArrowLLT
//   src/decomposition/blockArrowLLT.cpp:92-174     blockArrowLSolve_<Up> / blockArrowLSolve
//   src/decomposition/blockArrowLLT.cpp:176-277    blockArrowLTransposeSolve_<Up> / ...
//   src/structured/StructuredG.cpp:6-113           StructuredG (type tag + dispatch)
//
// The Eigen kernels those files call (llt_inplace, triangular solveInPlace, rankUpdate, gemm) are
// restated with ONE canonical operation order, the same family as gi_oracle.hpp:
//   chol(D)            left-looking, v_i = D(i,k) - dot4_{j<k}(L(i,j), L(k,j)); L(k,k) = sqrt(v_k);
//                      L(i,k) = v_i / L(k,k); fails when v_k <= 0 (Eigen: "if (x <= 0) return k").
//   B = S L^-T         row r of B: B(r,k) = (S(r,k) - dot4_{j<k}(B(r,j), L(k,j))) / L(k,k), k ascending
//                      (the same recurrence as a Cholesky row below the diagonal block). For the
//                      "up" arrow the reference solves L^-1 side[i] with side[i] = B^T: same numbers.
//   D -= B B^T         D(r,c) = D(r,c) - dot4_k(B(r,k), B(c,k)), r >= c (lower triangle only).
//   M_i -= B X         M(r,c) = M(r,c) - dot4_k(B(r,k), X(k,c));  transposed: dot4_k(B(k,r), X(k,c)).
//   L x = m, L^T x = m column-oriented substitution with true division by the diagonal and
//                      fma(-x_k, L(.,k), w) updates (as the dense trsv of gi_oracle.cpp).
// Every output element has a fixed order that does not depend on how outputs are distributed over
// threads, which is what lets the CUDA kernels match bit for bit.
// PARITY PINNING: checked against the reference's own tests for this path
// (tests/triBlockDiagLLTTest.cpp, tests/blockArrowLLTTest.cpp: agreement with the dense LLT and
// dense triangular solves to 1e-8, all (start, end) windows) in tests/test_decomp_oracle.py.
#pragma once

#include <cstddef>
#include <cstdint>
#include <vector>

namespace decomp_oracle
{

// A view on one block: column-major, leading dimension ld (Eigen::Ref<MatrixXd>).
struct Block
{
  double * p = nullptr;
  int rows = 0, cols = 0, ld = 0;
  double & operator()(int r, int c) const { return p[r + static_cast<std::ptrdiff_t>(c) * ld]; }
};

// structured::StructuredG::Type (include/jrl-qp/structured/StructuredG.h:17-22), same order.
enum Type : int
{
  TriBlockDiagonal = 0,
  BlockArrowUp = 1,
  BlockArrowDown = 2
};

bool triBlockDiagLLT(const std::vector<Block> & diag, const std::vector<Block> & subDiag);
void triBlockDiagLSolve(const std::vector<Block> & diag, const std::vector<Block> & subDiag, double * M, int ldm, int ncols, int start = 0);
void triBlockDiagLTransposeSolve(const std::vector<Block> & diag,
                                 const std::vector<Block> & subDiag,
                                 double * M,
                                 int ldm,
                                 int ncols,
                                 int end = -1);

bool blockArrowLLT(const std::vector<Block> & diag, const std::vector<Block> & side, bool up = false);
void blockArrowLSolve(const std::vector<Block> & diag,
                      const std::vector<Block> & side,
                      bool up,
                      double * M,
                      int ldm,
                      int ncols,
                      int start = 0,
                      int end = -1);
void blockArrowLTransposeSolve(const std::vector<Block> & diag,
                               const std::vector<Block> & side,
                               bool up,
                               double * M,
                               int ldm,
                               int ncols,
                               int start = 0,
                               int end = -1);

// structured::StructuredG restated (src/structured/StructuredG.cpp): type tag + dispatch.
class StructuredG
{
public:
  StructuredG() = default;
  StructuredG(Type t, const std::vector<Block> & diag, const std::vector<Block> & offDiag);
  Type type() const { return type_; }
  int nbVar() const { return nbVar_; }
  int nbVar(int i) const { return diag_[static_cast<size_t>(i)].cols; }
  bool lltInPlace();
  bool decomposed() const { return decomposed_; }
  void solveInPlaceLTranspose(double * v) const;
  void solveL(double * out, const double * in) const;
  // in = [0; v; 0] with nonzero rows [start, end) (internal::SingleNZSegmentVector overload)
  void solveL(double * out, const double * in, int start, int end) const;

private:
  Type type_ = TriBlockDiagonal;
  std::vector<Block> diag_, offDiag_;
  int nbVar_ = 0;
  bool decomposed_ = false;
};

} // namespace decomp_oracle
