// FP64 roofline probe: measures the achievable DFMA rate of the device with a dependent-free
// FMA loop (8 independent chains per thread, enough resident warps to saturate the FP64 pipe).
// MEASURED_PEAKS.json carries HBM and bf16 peaks only; the solver's roof is the FP64 vector pipe
// (SURVEY.md §8(d)), so bench.py takes its denominator from this probe.
#include "jrlqp_b200.h"

#include <cuda_runtime.h>

namespace
{

__global__ void __launch_bounds__(256) dfma_probe_kernel(double * out, int iters, double seed)
{
  double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 0.999999, c = 1e-9;
  for(int i = 0; i < iters; ++i)
  {
#pragma unroll
    for(int u = 0; u < 8; ++u)
    {
      a0 = fma(a0, m, c);
      a1 = fma(a1, m, c);
      a2 = fma(a2, m, c);
      a3 = fma(a3, m, c);
      a4 = fma(a4, m, c);
      a5 = fma(a5, m, c);
      a6 = fma(a6, m, c);
      a7 = fma(a7, m, c);
    }
  }
  double r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if(r == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = r; // keep the chains alive
}

} // namespace

extern "C" double jrlqp_measure_fp64_tflops(int32_t device, int32_t repeats)
{
  if(cudaSetDevice(device) != cudaSuccess) return -1.0;
  cudaDeviceProp prop;
  if(cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1.0;
  const int blocks = prop.multiProcessorCount * 8;
  const int threads = 256;
  const int iters = 4096;
  double * d = nullptr;
  if(cudaMalloc(&d, sizeof(double) * blocks * threads) != cudaSuccess) return -1.0;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0.0;
  if(repeats < 1) repeats = 1;
  for(int r = 0; r < repeats + 1; ++r)
  {
    cudaEventRecord(e0);
    dfma_probe_kernel<<<blocks, threads>>>(d, iters, 1.0 + r);
    cudaEventRecord(e1);
    if(cudaEventSynchronize(e1) != cudaSuccess) break;
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    double flops = 2.0 * 64.0 * iters * (double)blocks * threads;
    double tf = flops / (ms * 1e-3) / 1e12;
    if(r > 0 && tf > best) best = tf; // first launch is warm-up
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  return best > 0 ? best : -1.0;
}

// ---------------------------------------------------------------------------------------------
// Self-test of the short-latency exact primitives of fp64_exact.cuh against the stock IEEE
// operations, on pseudo-random operands (counter-based, so the test needs no input buffers).
// ---------------------------------------------------------------------------------------------
#include "fp64_exact.cuh"

namespace
{

__device__ __forceinline__ unsigned long long mix64(unsigned long long z)
{
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

// mantissa from the hash, exponent in [-span, span], random sign
__device__ __forceinline__ double rnd_double(unsigned long long h, int span)
{
  const unsigned long long mant = h & 0x000fffffffffffffull;
  const int ex = span ? (int)((h >> 52) % (unsigned)(2 * span + 1)) - span : 0;
  const unsigned long long bits = ((h >> 63) << 63) | ((unsigned long long)(1023 + ex) << 52) | mant;
  return __longlong_as_double((long long)bits);
}

// counts[0] quotients proven, [1] proven but != x / y (must stay 0), [2] unproven,
// [3] square roots != sqrt() (must stay 0), [4] reciprocal by-product further than 4 ulp from 1/sqrt
__global__ void arith_selftest_kernel(unsigned long long seed, int per_thread, int span, int rcp_ulps, unsigned long long * counts)
{
  unsigned long long c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0;
  const unsigned long long gid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  for(int it = 0; it < per_thread; ++it)
  {
    const unsigned long long k = mix64(seed + gid * 0x100000001b3ull + it);
    const double x = rnd_double(mix64(k), span), y = rnd_double(mix64(k ^ 0x5555555555555555ull), span);
    // reciprocal of y perturbed by up to rcp_ulps ulps, as the recurrences hand it over
    const double rc = 1.0 / y;
    const long long pert = (long long)(mix64(k + 7) % (unsigned)(2 * rcp_ulps + 1)) - rcp_ulps;
    const double r = __longlong_as_double(__double_as_longlong(rc) + pert);
    bool ok;
    const double qd = jrlqp::div_rcp(x, y, r, ok);
    const double qs = x / y;
    if(ok)
    {
      ++c0;
      if(__double_as_longlong(qd) != __double_as_longlong(qs)) ++c1;
    }
    else
      ++c2;
    // square root of 1 + t^2, |t| <= 1 (and of the raw mantissa in [1, 2))
    const double t = rnd_double(mix64(k + 13), 0) - 1.0; // [0, 1) with sign
    const double a = (it & 1) ? fma(t, t, 1.0) : fabs(rnd_double(mix64(k + 17), 0));
    double y1;
    const double sd = jrlqp::sqrt_rsqrt(a, y1);
    if(__double_as_longlong(sd) != __double_as_longlong(sqrt(a))) ++c3;
    const double yr = 1.0 / sqrt(a);
    if(fabs(y1 - yr) > 4.0 * 2.220446049250313e-16 * yr) ++c4;
  }
  atomicAdd(counts + 0, c0);
  atomicAdd(counts + 1, c1);
  atomicAdd(counts + 2, c2);
  atomicAdd(counts + 3, c3);
  atomicAdd(counts + 4, c4);
}

} // namespace

extern "C" int jrlqp_selftest_arith(int32_t device, int64_t samples, uint64_t seed, int32_t exponent_span, int32_t rcp_ulps, uint64_t * counts5)
{
  if(!counts5 || samples < 1 || exponent_span < 0 || exponent_span > 1000 || rcp_ulps < 0) return JRLQP_ERR_ARG;
  if(cudaSetDevice(device) != cudaSuccess) return JRLQP_ERR_CUDA;
  unsigned long long * d = nullptr;
  if(cudaMalloc(&d, 5 * sizeof(unsigned long long)) != cudaSuccess) return JRLQP_ERR_CUDA;
  cudaMemset(d, 0, 5 * sizeof(unsigned long long));
  const int threads = 256, blocks = 148 * 8, per_thread = (int)((samples + (long long)threads * blocks - 1) / ((long long)threads * blocks));
  arith_selftest_kernel<<<blocks, threads>>>(seed, per_thread, exponent_span, rcp_ulps, d);
  unsigned long long h[5];
  cudaError_t e = cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  cudaFree(d);
  if(e != cudaSuccess) return JRLQP_ERR_CUDA;
  for(int i = 0; i < 5; ++i) counts5[i] = h[i];
  return JRLQP_OK;
}
