// FP64 latency microbenchmark (one warp, dependent chains), run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false scripts/microbench.cu -o /tmp/mb && /tmp/mb
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
template<int OP> __global__ void chain(double x0, double y0, double * out, long long * cyc)
{
  double x = x0 + threadIdx.x * 1e-9, y = y0;
  __shared__ double sm[64];
  sm[threadIdx.x] = y0;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 16
  for(int i = 0; i < N; ++i)
  {
    if(OP == 0) x = fma(x, y, y);
    if(OP == 1) x = x * y;
    if(OP == 2) x = x + y;
    if(OP == 3) x = y / x;
    if(OP == 4) x = sqrt(fma(x, x, 1.0)) * 0.5;
    if(OP == 5) x = 1.0 / x;
    if(OP == 6) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);
    if(OP == 7) { int k = ((int)x) & 31; x = sm[k] + 1.0; } // LDS + DADD dependent
    if(OP == 8) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r + 1.5; }
    if(OP == 9) { double t = y / x; double u = sqrt(fma(t, t, 1.0)); x = x * u; } // one Givens link (kind 3)
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if(threadIdx.x == 0) *cyc = t1 - t0;
}
template<int OP> void run(const char * name, double x0, double y0, double sub = 0)
{
  double * out; long long * cyc, h;
  cudaMalloc(&out, 256); cudaMalloc(&cyc, 8);
  chain<OP><<<1, 32>>>(x0, y0, out, cyc);
  chain<OP><<<1, 32>>>(x0, y0, out, cyc);
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-34s %7.1f cycles/iter\n", name, (double)h / N - sub);
  cudaFree(out); cudaFree(cyc);
}
int main()
{
  run<0>("DFMA dependent", 0.5, 0.999);
  run<1>("DMUL dependent", 1.0, 1.0000001);
  run<2>("DADD dependent", 1.0, 1e-9);
  run<3>("DDIV (y/x) dependent", 1.5, 2.0);
  run<4>("sqrt(fma(x,x,1))*0.5 dependent", 1.5, 2.0);
  run<5>("1.0/x dependent", 1.5, 2.0);
  run<6>("SHFL f64 dependent", 1.5, 2.0);
  run<7>("LDS + F2I + DADD dependent", 1.5, 2.0);
  run<8>("MUFU.RCP64H + DADD dependent", 1.5, 2.0);
  run<9>("Givens link: div, fma, sqrt, mul", 1.5, 0.01);
  return 0;
}
