// FP64 chain latency under contention: k single-warp CTAs per SM (148*k CTAs), each running a dependent chain.
#include <cstdio>
#include <cuda_runtime.h>
#define N 2048
template<int OP> __global__ void chain(double x0, double y0, double * out, long long * cyc)
{
  double x = x0 + threadIdx.x * 1e-9 + blockIdx.x * 1e-7, y = y0;
  long long t0 = clock64();
#pragma unroll 8
  for(int i = 0; i < N; ++i)
  {
    if(OP == 0) x = fma(x, y, y);
    if(OP == 9) { double t = y / x; double u = sqrt(fma(t, t, 1.0)); x = x * u; }
  }
  long long t1 = clock64();
  out[blockIdx.x * 32 + threadIdx.x] = x;
  if(threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template<int OP> void run(const char * name, double x0, double y0, int k, int threads)
{
  int blocks = 148 * k;
  double * out; long long * cyc;
  cudaMalloc(&out, blocks * threads * 8); cudaMalloc(&cyc, 8 * blocks);
  chain<OP><<<blocks, threads>>>(x0, y0, out, cyc);
  chain<OP><<<blocks, threads>>>(x0, y0, out, cyc);
  long long * h = new long long[blocks];
  cudaMemcpy(h, cyc, 8 * blocks, cudaMemcpyDeviceToHost);
  double s = 0; for(int i = 0; i < blocks; ++i) s += h[i];
  printf("%-28s CTAs/SM %2d x %3d thr: %7.1f cycles/iter (mean over CTAs)\n", name, k, threads, s / blocks / N);
  cudaFree(out); cudaFree(cyc); delete[] h;
}
int main()
{
  for(int k : {1, 2, 4, 6, 8, 12, 16, 24}) run<0>("DFMA dependent", 0.5, 0.999, k, 32);
  for(int k : {1, 2, 4, 6, 8, 12, 16, 24}) run<9>("Givens link (stock div/sqrt)", 1.5, 0.01, k, 32);
  for(int k : {1, 2, 4, 6}) run<9>("Givens link (stock div/sqrt)", 1.5, 0.01, k, 64);
  return 0;
}
