// TEST INFRASTRUCTURE — NOT PRODUCT CODE. See decomp_oracle.hpp for scope, the reference files
// restated and the canonical arithmetic order.
#include "decomp_oracle.hpp"

#include "gi_oracle.hpp"

#include <algorithm>
#include <cassert>
#include <cmath>

namespace decomp_oracle
{

namespace
{

using gi_oracle::dot4;

// strided view: element (r, c) at p[r * rs + c * cs]; lets "side[i].transpose()" be a view.
struct View
{
  double * p;
  int rows, cols;
  std::ptrdiff_t rs, cs;
  double & operator()(int r, int c) const { return p[r * rs + c * cs]; }
  View t() const { return {p, cols, rows, cs, rs}; }
  View sub(int r0, int c0, int nr, int nc) const { return {p + r0 * rs + c0 * cs, nr, nc, rs, cs}; }
};

View view(const Block & b)
{
  return {b.p, b.rows, b.cols, 1, b.ld};
}

// Eigen::internal::llt_inplace<double, Lower>::blocked restated (canonical left-looking order).
bool chol(const View & D)
{
  const int n = D.rows;
  std::vector<double> w(static_cast<size_t>(n));
  for(int k = 0; k < n; ++k)
  {
    for(int i = k; i < n; ++i) w[static_cast<size_t>(i)] = D(i, k) - dot4(k, &D(i, 0), D.cs, &D(k, 0), D.cs);
    if(w[static_cast<size_t>(k)] <= 0.0) return false;
    double lkk = std::sqrt(w[static_cast<size_t>(k)]);
    D(k, k) = lkk;
    for(int i = k + 1; i < n; ++i) D(i, k) = w[static_cast<size_t>(i)] / lkk;
  }
  return true;
}

// B = B L^-T (Li.transpose().solveInPlace<OnTheRight>(S), or Li.solveInPlace<OnTheLeft>(side) seen
// through B = side^T): the rows of B continue the Cholesky recurrence below the diagonal block.
void trsmRightLT(const View & B, const View & L)
{
  const int n = L.rows;
  assert(B.cols == n);
  for(int r = 0; r < B.rows; ++r)
    for(int k = 0; k < n; ++k) B(r, k) = (B(r, k) - dot4(k, &B(r, 0), B.cs, &L(k, 0), L.cs)) / L(k, k);
}

// D.selfadjointView<Lower>().rankUpdate(B, -1)
void syrkSub(const View & D, const View & B)
{
  const int n = D.rows;
  assert(B.rows == n);
  for(int c = 0; c < n; ++c)
    for(int r = c; r < n; ++r) D(r, c) = D(r, c) - dot4(B.cols, &B(r, 0), B.cs, &B(c, 0), B.cs);
}

// M.noalias() -= B * X   (M: B.rows x ncols, X: B.cols x ncols)
void gemmSub(const View & M, const View & B, const View & X)
{
  assert(M.rows == B.rows && X.rows == B.cols && M.cols == X.cols);
  for(int c = 0; c < M.cols; ++c)
    for(int r = 0; r < M.rows; ++r) M(r, c) = M(r, c) - dot4(B.cols, &B(r, 0), B.cs, &X(0, c), X.rs);
}

// L.triangularView<Lower>().solveInPlace(M)
void solveLower(const View & L, const View & M)
{
  const int n = L.rows;
  assert(M.rows == n);
  for(int c = 0; c < M.cols; ++c)
    for(int k = 0; k < n; ++k)
    {
      double xk = M(k, c) / L(k, k);
      M(k, c) = xk;
      for(int i = k + 1; i < n; ++i) M(i, c) = std::fma(-xk, L(i, k), M(i, c));
    }
}

// L.triangularView<Lower>().transpose().solveInPlace(M)
void solveLowerT(const View & L, const View & M)
{
  const int n = L.rows;
  assert(M.rows == n);
  for(int c = 0; c < M.cols; ++c)
    for(int k = n - 1; k >= 0; --k)
    {
      double xk = M(k, c) / L(k, k);
      M(k, c) = xk;
      for(int i = 0; i < k; ++i) M(i, c) = std::fma(-xk, L(k, i), M(i, c));
    }
}

View rowsOf(double * M, int ldm, int ncols, int r0, int nr)
{
  return {M + r0, nr, ncols, 1, ldm};
}

// helper struct get<Up> of src/decomposition/blockArrowLLT.cpp:11-48
View getD(const std::vector<Block> & diag, int i, bool up)
{
  return view(up ? diag[static_cast<size_t>(i + 1) % diag.size()] : diag[static_cast<size_t>(i)]);
}
View getB(const std::vector<Block> & side, int i, bool up)
{
  View s = view(side[static_cast<size_t>(i)]);
  return up ? s.t() : s;
}

int totalRows(const std::vector<Block> & diag)
{
  int n = 0;
  for(const auto & d : diag) n += d.rows;
  return n;
}

} // namespace

// src/decomposition/triBlockDiagLLT.cpp:9-36
bool triBlockDiagLLT(const std::vector<Block> & diag, const std::vector<Block> & subDiag)
{
  assert(diag.size() == subDiag.size() + 1);
  size_t b = diag.size();
  for(size_t i = 0; i + 1 < b; ++i)
  {
    View Di = view(diag[i]);
    if(!chol(Di)) return false; // Li = chol(Di)
    trsmRightLT(view(subDiag[i]), Di); // Si = Si Li^-T
    syrkSub(view(diag[i + 1]), view(subDiag[i])); // D[i+1] -= Si Si^T
  }
  return chol(view(diag.back()));
}

// src/decomposition/triBlockDiagLLT.cpp:38-98
void triBlockDiagLSolve(const std::vector<Block> & diag, const std::vector<Block> & subDiag, double * M, int ldm, int ncols, int start)
{
  assert(diag.size() == subDiag.size() + 1);
  int n = 0, l = 0, li = 0;
  bool zero = true;
  for(size_t i = 0; i < diag.size(); ++i)
  {
    View Di = view(diag[i]);
    int ni = Di.rows;
    if(n + ni >= start)
    {
      if(zero)
      {
        int r = n + ni - start;
        solveLower(Di.sub(ni - r, ni - r, r, r), rowsOf(M, ldm, ncols, start, r));
        zero = false;
      }
      else
      {
        View Mi = rowsOf(M, ldm, ncols, n, ni);
        gemmSub(Mi, view(subDiag[i - 1]), rowsOf(M, ldm, ncols, l, li));
        solveLower(Di, Mi);
      }
    }
    l = n;
    li = ni;
    n += ni;
  }
}

// src/decomposition/triBlockDiagLLT.cpp:100-158
void triBlockDiagLTransposeSolve(const std::vector<Block> & diag, const std::vector<Block> & subDiag, double * M, int ldm, int ncols, int end)
{
  int n = totalRows(diag);
  int l = 0, li = 0;
  bool zero = true;
  if(end < 0) end = n;
  for(int i = static_cast<int>(diag.size()) - 1; i >= 0; --i)
  {
    View Di = view(diag[static_cast<size_t>(i)]);
    int ni = Di.rows;
    if(n - ni < end)
    {
      if(zero)
      {
        int r = end - n + ni;
        solveLowerT(Di.sub(0, 0, r, r), rowsOf(M, ldm, ncols, n - ni, r));
        zero = false;
      }
      else
      {
        View Mi = rowsOf(M, ldm, ncols, n - ni, ni);
        gemmSub(Mi, view(subDiag[static_cast<size_t>(i)]).t(), rowsOf(M, ldm, ncols, l - li, li));
        solveLowerT(Di, Mi);
      }
    }
    l = n;
    li = ni;
    n -= ni;
  }
}

// src/decomposition/blockArrowLLT.cpp:52-90
bool blockArrowLLT(const std::vector<Block> & diag, const std::vector<Block> & side, bool up)
{
  assert(diag.size() == side.size() + 1);
  int b = static_cast<int>(diag.size());
  View Db = getD(diag, b - 1, up);
  for(int i = 0; i < b - 1; ++i)
  {
    View Di = getD(diag, i, up);
    if(!chol(Di)) return false;
    View Bi = getB(side, i, up);
    trsmRightLT(Bi, Di); // Bi = Bi Li^-T   (up: side[i] = Li^-1 side[i])
    syrkSub(Db, Bi); // Db -= Bi Bi^T
  }
  return chol(Db);
}

namespace
{

// src/decomposition/blockArrowLLT.cpp:92-152
void arrowLSolve_(const std::vector<Block> & diag, const std::vector<Block> & side, bool up, double * M, int ldm, int ncols, int start, int end)
{
  int b = static_cast<int>(diag.size());
  int n = 0;
  // The reference writes "M.bottomRows(diag.back().rows())" (blockArrowLLT.cpp:142), which for Up is
  // the size of the wrong block unless diag.front() and diag.back() have the same size (they do in
  // its tests); the rows meant are those of the last block of the permuted system, used here.
  int nbLast = getD(diag, b - 1, up).rows;
  int total = totalRows(diag);
  for(int i = 0; i < b - 1; ++i)
  {
    View Di = getD(diag, i, up);
    int ni = Di.rows;
    int s = std::max(start - n, 0);
    if(ni < s || end <= n)
    {
      n += ni;
      continue;
    }
    View Mi = rowsOf(M, ldm, ncols, n + s, ni - s);
    solveLower(Di.sub(s, s, ni - s, ni - s), Mi);
    View Bi = getB(side, i, up);
    gemmSub(rowsOf(M, ldm, ncols, total - nbLast, nbLast), Bi.sub(0, s, Bi.rows, ni - s), Mi);
    n += ni;
  }
  View Db = getD(diag, b - 1, up);
  solveLower(Db, rowsOf(M, ldm, ncols, n, Db.rows));
}

// src/decomposition/blockArrowLLT.cpp:176-252
void arrowLTSolve_(const std::vector<Block> & diag, const std::vector<Block> & side, bool up, double * M, int ldm, int ncols, int start, int end)
{
  int b = static_cast<int>(diag.size());
  int s = totalRows(diag);
  bool zero = false;
  View Db = getD(diag, b - 1, up);
  int nb = Db.rows;
  View Mb = rowsOf(M, ldm, ncols, s - nb, nb);
  if(end > s - nb)
  {
    int r = end - s + nb;
    solveLowerT(Db.sub(0, 0, r, r), Mb.sub(0, 0, r, ncols));
  }
  else
    zero = true;

  int n = 0;
  for(int i = 0; i < b - 1; ++i)
  {
    View Di = getD(diag, i, up);
    int ni = Di.rows;
    View Mi = rowsOf(M, ldm, ncols, n, ni);
    if(zero)
    {
      if(start >= n + ni)
      {
        n += ni;
        continue;
      }
    }
    else
      gemmSub(Mi, getB(side, i, up).t(), Mb);

    if(end >= n)
    {
      if(end >= n + ni)
        solveLowerT(Di, Mi);
      else
      {
        int r = end - n;
        solveLowerT(Di.sub(0, 0, r, r), Mi.sub(0, 0, r, ncols));
      }
    }
    n += ni;
  }
}

// v.topRows(s) = tmp.bottomRows(s); v.bottomRows(n0) = tmp.topRows(n0)  (blockArrowLLT.cpp:165-169)
void rotateUp(double * M, int ldm, int ncols, int total, int n0)
{
  std::vector<double> tmp(static_cast<size_t>(total));
  int s = total - n0;
  for(int c = 0; c < ncols; ++c)
  {
    double * v = M + static_cast<std::ptrdiff_t>(c) * ldm;
    std::copy(v, v + total, tmp.begin());
    for(int i = 0; i < s; ++i) v[i] = tmp[static_cast<size_t>(n0 + i)];
    for(int i = 0; i < n0; ++i) v[s + i] = tmp[static_cast<size_t>(i)];
  }
}

// v.bottomRows(s) = tmp.topRows(s); v.topRows(n0) = tmp.bottomRows(n0)  (blockArrowLLT.cpp:266-270)
void rotateDown(double * M, int ldm, int ncols, int total, int n0)
{
  std::vector<double> tmp(static_cast<size_t>(total));
  int s = total - n0;
  for(int c = 0; c < ncols; ++c)
  {
    double * v = M + static_cast<std::ptrdiff_t>(c) * ldm;
    std::copy(v, v + total, tmp.begin());
    for(int i = 0; i < s; ++i) v[n0 + i] = tmp[static_cast<size_t>(i)];
    for(int i = 0; i < n0; ++i) v[i] = tmp[static_cast<size_t>(s + i)];
  }
}

} // namespace

// src/decomposition/blockArrowLLT.cpp:154-174
void blockArrowLSolve(const std::vector<Block> & diag, const std::vector<Block> & side, bool up, double * M, int ldm, int ncols, int start, int end)
{
  int total = totalRows(diag);
  if(end < 0) end = total;
  if(up)
  {
    int n0 = diag.front().rows;
    rotateUp(M, ldm, ncols, total, n0);
    arrowLSolve_(diag, side, true, M, ldm, ncols, std::max(0, start - n0), std::max(0, end - n0));
  }
  else
    arrowLSolve_(diag, side, false, M, ldm, ncols, start, end);
}

// src/decomposition/blockArrowLLT.cpp:254-277
void blockArrowLTransposeSolve(const std::vector<Block> & diag,
                               const std::vector<Block> & side,
                               bool up,
                               double * M,
                               int ldm,
                               int ncols,
                               int start,
                               int end)
{
  int total = totalRows(diag);
  if(end < 0) end = total;
  if(up)
  {
    arrowLTSolve_(diag, side, true, M, ldm, ncols, start, end);
    rotateDown(M, ldm, ncols, total, diag.front().rows);
  }
  else
    arrowLTSolve_(diag, side, false, M, ldm, ncols, start, end);
}

// ---------------------------------------------------------------------------
// structured::StructuredG (src/structured/StructuredG.cpp)
// ---------------------------------------------------------------------------

StructuredG::StructuredG(Type t, const std::vector<Block> & diag, const std::vector<Block> & offDiag)
: type_(t), diag_(diag), offDiag_(offDiag)
{
  nbVar_ = totalRows(diag);
}

bool StructuredG::lltInPlace()
{
  bool done = false;
  switch(type_)
  {
    case TriBlockDiagonal:
      done = triBlockDiagLLT(diag_, offDiag_);
      break;
    case BlockArrowUp:
      done = blockArrowLLT(diag_, offDiag_, true);
      break;
    case BlockArrowDown:
      done = blockArrowLLT(diag_, offDiag_, false);
      break;
  }
  decomposed_ = done;
  return done;
}

void StructuredG::solveInPlaceLTranspose(double * v) const
{
  switch(type_)
  {
    case TriBlockDiagonal:
      triBlockDiagLTransposeSolve(diag_, offDiag_, v, nbVar_, 1);
      break;
    case BlockArrowUp:
      blockArrowLTransposeSolve(diag_, offDiag_, true, v, nbVar_, 1);
      break;
    case BlockArrowDown:
      blockArrowLTransposeSolve(diag_, offDiag_, false, v, nbVar_, 1);
      break;
  }
}

void StructuredG::solveL(double * out, const double * in) const
{
  solveL(out, in, 0, -1);
}

void StructuredG::solveL(double * out, const double * in, int start, int end) const
{
  if(out != in) std::copy(in, in + nbVar_, out);
  switch(type_)
  {
    case TriBlockDiagonal:
      triBlockDiagLSolve(diag_, offDiag_, out, nbVar_, 1, start);
      break;
    case BlockArrowUp:
      blockArrowLSolve(diag_, offDiag_, true, out, nbVar_, 1, start, end);
      break;
    case BlockArrowDown:
      blockArrowLSolve(diag_, offDiag_, false, out, nbVar_, 1, start, end);
      break;
  }
}

} // namespace decomp_oracle
