// Host-side state of a jrlqp_structured handle (include/jrlqp_b200.h), shared by structured.cu (the
// decomposition / solve entry points) and blockgi.cu (the structured solver built on them).
#pragma once

#include "structured.cuh"

#include "jrlqp_b200.h"

#include <algorithm>
#include <string>
#include <vector>

struct jrlqp_structured
{
  int type = 0, b = 0, n = 0, nmax = 0;
  long long capacity = 0;
  int device = 0;
  int threads = 32;
  int num_sms = 0;
  int llt_smem = 0, solve_smem = 0, llt_occ = 0, solve_occ = 0;
  long long touched = 0;
  std::vector<int> size, dld, old, start;
  std::vector<long long> doff, ooff;
  long long min_stride = 0; // one past the last element any block touches
  // small-tile kernel (structured_small.cuh): tile size if the structure is a tri-block-diagonal chain of uniform dense
  // tiles of 8, 12 or 16 rows at even offsets, else 0; kernel_mode: 0 automatic, 1 general, 2 small tiles, 3 small tiles + TMA
  int small_nb = 0;
  int kernel_mode = 0;
  // device copies of the descriptor
  int *d_size = nullptr, *d_dld = nullptr, *d_old = nullptr, *d_start = nullptr;
  long long *d_doff = nullptr, *d_ooff = nullptr;
  // staging for the host entry points
  double * d_data = nullptr;
  long long d_data_elems = 0;
  double * d_M = nullptr;
  long long d_M_elems = 0;
  int * d_ok = nullptr;
  cudaStream_t stream = nullptr;
  std::string err;

  bool check(cudaError_t e, const char * what)
  {
    if(e == cudaSuccess) return true;
    err = std::string(what) + ": " + cudaGetErrorString(e);
    return false;
  }
};

#define SCK(call)                                       \
  do                                                    \
  {                                                     \
    if(!s->check((call), #call)) return JRLQP_ERR_CUDA; \
  } while(0)

namespace jrlqp
{

inline StructParams base_params(const jrlqp_structured * s)
{
  StructParams p{};
  p.type = s->type;
  p.b = s->b;
  p.n = s->n;
  p.nmax = s->nmax;
  p.size = s->d_size;
  p.doff = s->d_doff;
  p.dld = s->d_dld;
  p.ooff = s->d_ooff;
  p.old = s->d_old;
  p.start = s->d_start;
  return p;
}

template<class T>
inline cudaError_t upload(T *& d, const std::vector<T> & h)
{
  cudaError_t e = cudaMalloc(&d, sizeof(T) * std::max<size_t>(h.size(), 1));
  if(e != cudaSuccess) return e;
  if(h.empty()) return cudaSuccess;
  return cudaMemcpy(d, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice);
}

} // namespace jrlqp
