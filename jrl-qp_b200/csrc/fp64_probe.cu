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
