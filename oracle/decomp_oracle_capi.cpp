// TEST INFRASTRUCTURE — NOT PRODUCT CODE. C entry points over decomp_oracle.cpp for ctypes
// (tests/, bench.py's cpu_baseline / --impl reference legs). The structure descriptor is the one
// of include/jrlqp_b200.h (jrlqp_structure): block sizes, and for every diagonal / off-diagonal
// block an element offset from the instance base and a leading dimension.
#include "decomp_oracle.hpp"

#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

using namespace decomp_oracle;

namespace
{

struct Desc
{
  int type, b;
  const int * size;
  const long * doff;
  const int * dld;
  const long * ooff;
  const int * old;
};

void views(const Desc & d, double * base, std::vector<Block> & diag, std::vector<Block> & off)
{
  diag.resize(static_cast<size_t>(d.b));
  off.resize(static_cast<size_t>(d.b - 1));
  for(int i = 0; i < d.b; ++i) diag[static_cast<size_t>(i)] = {base + d.doff[i], d.size[i], d.size[i], d.dld[i]};
  for(int i = 0; i + 1 < d.b; ++i)
  {
    int rows, cols;
    if(d.type == TriBlockDiagonal)
    {
      rows = d.size[i + 1];
      cols = d.size[i];
    }
    else if(d.type == BlockArrowDown)
    {
      rows = d.size[d.b - 1];
      cols = d.size[i];
    }
    else
    {
      rows = d.size[i + 1];
      cols = d.size[0];
    }
    off[static_cast<size_t>(i)] = {base + d.ooff[i], rows, cols, d.old[i]};
  }
}

template<class F>
void parallelFor(long batch, int nthreads, F f)
{
  if(nthreads < 1) nthreads = 1;
  if(nthreads > batch) nthreads = static_cast<int>(std::max<long>(1, batch));
  std::atomic<long> next(0);
  auto worker = [&]()
  {
    for(;;)
    {
      long b0 = next.fetch_add(8);
      if(b0 >= batch) break;
      for(long k = b0; k < std::min(batch, b0 + 8); ++k) f(k);
    }
  };
  if(nthreads == 1)
    worker();
  else
  {
    std::vector<std::thread> th;
    for(int t = 0; t < nthreads; ++t) th.emplace_back(worker);
    for(auto & t : th) t.join();
  }
}

} // namespace

extern "C"
{

/** StructuredG::lltInPlace over `batch` instances laid out `stride` elements apart, in place.
 * ok[k] = 1 when instance k was decomposed (0: a diagonal block was not positive definite; the
 * failing block and the blocks after it are then left partially updated, as in the reference). */
int decomp_oracle_llt(int type,
                      int b,
                      const int * size,
                      const long * doff,
                      const int * dld,
                      const long * ooff,
                      const int * old,
                      double * data,
                      long stride,
                      long batch,
                      int * ok,
                      int nthreads)
{
  Desc d{type, b, size, doff, dld, ooff, old};
  parallelFor(batch, nthreads,
              [&](long k)
              {
                std::vector<Block> diag, off;
                views(d, data + k * stride, diag, off);
                bool r = type == TriBlockDiagonal ? triBlockDiagLLT(diag, off) : blockArrowLLT(diag, off, type == BlockArrowUp);
                if(ok) ok[k] = r ? 1 : 0;
              });
  return 0;
}

/** StructuredG::solveL (transpose = 0) / solveInPlaceLTranspose (transpose = 1) in place on
 * M (n x ncols, leading dimension ldm, `mstride` elements between instances), with the reference's
 * start / end hints (end < 0: none). `data` holds the factor produced by decomp_oracle_llt. */
int decomp_oracle_solve(int type,
                        int b,
                        const int * size,
                        const long * doff,
                        const int * dld,
                        const long * ooff,
                        const int * old,
                        const double * data,
                        long stride,
                        double * M,
                        int ldm,
                        int ncols,
                        long mstride,
                        long batch,
                        int transpose,
                        int start,
                        int end,
                        int nthreads)
{
  Desc d{type, b, size, doff, dld, ooff, old};
  parallelFor(batch, nthreads,
              [&](long k)
              {
                std::vector<Block> diag, off;
                views(d, const_cast<double *>(data) + k * stride, diag, off);
                double * Mk = M + k * mstride;
                if(type == TriBlockDiagonal)
                {
                  if(transpose)
                    triBlockDiagLTransposeSolve(diag, off, Mk, ldm, ncols, end);
                  else
                    triBlockDiagLSolve(diag, off, Mk, ldm, ncols, start);
                }
                else
                {
                  if(transpose)
                    blockArrowLTransposeSolve(diag, off, type == BlockArrowUp, Mk, ldm, ncols, start, end);
                  else
                    blockArrowLSolve(diag, off, type == BlockArrowUp, Mk, ldm, ncols, start, end);
                }
              });
  return 0;
}

} // extern "C"
