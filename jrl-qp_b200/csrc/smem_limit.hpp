// cudaFuncAttributeMaxDynamicSharedMemorySize is state of the (function, device) pair, shared by every handle of the
// process that launches that function: a handle that sets it to ITS size would lower it under the feet of another live
// handle that needs more (two solvers of different n in the same kernel instantiation, two structured handles ...).
// The limit is therefore only ever raised: one process-wide table of the largest size asked for so far.
#pragma once

#include <cuda_runtime.h>

#include <map>
#include <mutex>
#include <utility>

namespace jrlqp
{

inline cudaError_t raise_smem_limit(const void * fn, int bytes)
{
  static std::mutex mu;
  static std::map<std::pair<const void *, int>, int> current;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if(e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lk(mu);
  int & cur = current[std::make_pair(fn, dev)];
  if(bytes <= cur) return cudaSuccess;
  e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if(e == cudaSuccess) cur = bytes;
  return e;
}

template<typename F>
inline cudaError_t raise_smem_limit(F * fn, int bytes)
{
  return raise_smem_limit(reinterpret_cast<const void *>(fn), bytes);
}

} // namespace jrlqp
