// The FP64 tensor-core experiment north_star asks for ("DMMA only for the dense initial J = L^-T and block-Cholesky
// products at n >= 64, and only where ncu shows it beating the FP64 pipe"): jrlqp_probe_dmma measures, on the GEMM-shaped
// piece those two steps consist of (a 128 x 128 x 64 product per CTA, operands in shared memory, one 4-warp CTA per
// product — the shape and residency of the n = 128 kernel),
//   [0] GFLOP/s of a register-blocked FP64-pipe kernel (8 x 8 outputs per thread, 64 DFMA per 16 LDS),
//   [1] GFLOP/s of an mma.sync.m8n8k4.f64 (DMMA) kernel (32 x 64 outputs per warp, 32 DMMA per 12 LDS),
// and what decides whether DMMA can be used at all under the bit-exact contract of this library:
//   [2] fraction of DMMA outputs that equal the SEQUENTIAL chain fma(a3,b3, fma(a2,b2, fma(a1,b1, fma(a0,b0,c)))),
//   [3] fraction that equal the pairwise order ((a0 b0 + a1 b1) + (a2 b2 + a3 b3)) + c (fused products),
//   [4] fraction of length-64 inner products, evaluated with DMMA as FOUR accumulator tiles fed with k = j, j+4, j+8, j+12
//       (the interleaving of the canonical dot4) and combined as (c0 + c1) + (c2 + c3), that equal dot4 bit
//       for bit — if [2] is 1 this must be 1 too, and DMMA can replace the FP64 pipe without touching the oracle,
//   [5] the FP64-pipe kernel and the DMMA kernel agree bit for bit when both use sequential-k accumulation (0/1).
// Results: profiles/r02*_dmma_probe.txt; discussion in DESIGN.md §4.9.
#include "jrlqp_b200.h"

#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

namespace
{

constexpr int PM = 128, PN = 128, PK = 64;

__device__ __forceinline__ void dmma(double & d0, double & d1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__device__ __forceinline__ unsigned long long mix(unsigned long long x)
{
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__device__ __forceinline__ double urand(unsigned long long s)
{
  return (double)(mix(s) >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0;
}

// As[k][m] and Bs[k][n] (m, n contiguous): C(m, n) = sum_k A(m, k) B(k, n), sequential in k, one accumulator per output
__global__ void __launch_bounds__(128) gemm_fp64_pipe(double * C, int reps, unsigned long long seed)
{
  extern __shared__ __align__(16) double sm[];
  double * As = sm;
  double * Bs = sm + PK * PM;
  for(int i = threadIdx.x; i < PK * PM; i += blockDim.x) As[i] = urand(seed + blockIdx.x * 1000003ull + i);
  for(int i = threadIdx.x; i < PK * PN; i += blockDim.x) Bs[i] = urand(seed + 77 + blockIdx.x * 1000003ull + i);
  __syncthreads();
  // thread (tm, tn): rows 8 tm .. 8 tm + 7, two column groups 8 tn .. and 64 + 8 tn ..
  const int tm = threadIdx.x & 15, tn = threadIdx.x >> 4;
#pragma unroll 1
  for(int pass = 0; pass < 2; ++pass)
  {
    double acc[8][8];
#pragma unroll
    for(int i = 0; i < 8; ++i)
#pragma unroll
      for(int j = 0; j < 8; ++j) acc[i][j] = 0.0;
    const double * ap = As + 8 * tm;
    const double * bp = Bs + 64 * pass + 8 * tn;
    // the accumulators run on over the repetitions (every FMA of every repetition is live)
#pragma unroll 1
    for(int rep = 0; rep < reps; ++rep)
    {
#pragma unroll 2
      for(int k = 0; k < PK; ++k)
      {
        double a[8], b[8];
#pragma unroll
        for(int i = 0; i < 8; i += 2)
        {
          const double2 t = *reinterpret_cast<const double2 *>(ap + k * PM + i);
          a[i] = t.x;
          a[i + 1] = t.y;
          const double2 u = *reinterpret_cast<const double2 *>(bp + k * PN + i);
          b[i] = u.x;
          b[i + 1] = u.y;
        }
#pragma unroll
        for(int i = 0; i < 8; ++i)
#pragma unroll
          for(int j = 0; j < 8; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
      }
    }
#pragma unroll
    for(int i = 0; i < 8; ++i)
#pragma unroll
      for(int j = 0; j < 8; ++j) C[((size_t)blockIdx.x * PM + 8 * tm + i) * PN + 64 * pass + 8 * tn + j] = acc[i][j];
  }
}

__global__ void __launch_bounds__(128) gemm_dmma(double * C, int reps, unsigned long long seed)
{
  extern __shared__ __align__(16) double sm[];
  double * As = sm;
  double * Bs = sm + PK * PM;
  for(int i = threadIdx.x; i < PK * PM; i += blockDim.x) As[i] = urand(seed + blockIdx.x * 1000003ull + i);
  for(int i = threadIdx.x; i < PK * PN; i += blockDim.x) Bs[i] = urand(seed + 77 + blockIdx.x * 1000003ull + i);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3; // fragment coordinates: A(row g, k t), B(k t, col g), C(row g, cols 2 t, 2 t + 1)
#pragma unroll 1
  for(int pass = 0; pass < 2; ++pass)
  {
    // warp tile: rows 32 warp .. + 31 (4 tiles), columns 64 pass .. + 63 (8 tiles)
    double acc[4][8][2];
#pragma unroll
    for(int i = 0; i < 4; ++i)
#pragma unroll
      for(int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll 1
    for(int rep = 0; rep < reps; ++rep)
    {
#pragma unroll 2
      for(int k0 = 0; k0 < PK; k0 += 4)
      {
        double a[4], b[8];
#pragma unroll
        for(int i = 0; i < 4; ++i) a[i] = As[(k0 + t) * PM + 32 * warp + 8 * i + g];
#pragma unroll
        for(int j = 0; j < 8; ++j) b[j] = Bs[(k0 + t) * PN + 64 * pass + 8 * j + g];
#pragma unroll
        for(int i = 0; i < 4; ++i)
#pragma unroll
          for(int j = 0; j < 8; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
      }
    }
#pragma unroll
    for(int i = 0; i < 4; ++i)
#pragma unroll
      for(int j = 0; j < 8; ++j)
      {
        double * c = C + ((size_t)blockIdx.x * PM + 32 * warp + 8 * i + g) * PN + 64 * pass + 8 * j + 2 * t;
        c[0] = acc[i][j][0];
        c[1] = acc[i][j][1];
      }
  }
}

// accumulation order of ONE DMMA, and the dot4 emulation
__global__ void dmma_order(unsigned long long seed, int trials, unsigned long long * counts)
{
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  unsigned long long seq = 0, pair = 0, tot = 0, d4 = 0, d4tot = 0;
  for(int tr = 0; tr < trials; ++tr)
  {
    const unsigned long long s = seed + ((unsigned long long)blockIdx.x * trials + tr) * 4096ull;
    // A (8 x 4), B (4 x 8), C (8 x 8): element values are functions of their coordinates so that every lane can rebuild them
    auto A = [&](int r, int k) { return urand(s + 1 + r * 4 + k) * (1.0 + 3.0 * ((r + k) & 1)); };
    auto Bv = [&](int k, int c) { return urand(s + 100 + k * 8 + c); };
    auto Cv = [&](int r, int c) { return urand(s + 200 + r * 8 + c) * 4.0; };
    double d0 = Cv(g, 2 * t), d1 = Cv(g, 2 * t + 1);
    dmma(d0, d1, A(g, t), Bv(t, g));
#pragma unroll
    for(int e = 0; e < 2; ++e)
    {
      const int c = 2 * t + e;
      const double got = e ? d1 : d0;
      double r = Cv(g, c);
      for(int k = 0; k < 4; ++k) r = fma(A(g, k), Bv(k, c), r);
      const double p = (fma(A(g, 0), Bv(0, c), A(g, 1) * Bv(1, c)) + fma(A(g, 2), Bv(2, c), A(g, 3) * Bv(3, c))) + Cv(g, c);
      seq += got == r;
      pair += got == p;
      ++tot;
    }
    // dot4 of length 64 per output (row g of X, column c of Y): four accumulator tiles fed with k = j, j+4, j+8, j+12 per DMMA
    auto X = [&](int r, int k) { return urand(s + 1000 + r * 64 + k); };
    auto Y = [&](int k, int c) { return urand(s + 2000 + k * 8 + c); };
    double acc[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
    for(int k0 = 0; k0 < 64; k0 += 16)
#pragma unroll
      for(int j = 0; j < 4; ++j) dmma(acc[j][0], acc[j][1], X(g, k0 + j + 4 * t), Y(k0 + j + 4 * t, g));
#pragma unroll
    for(int e = 0; e < 2; ++e)
    {
      const int c = 2 * t + e;
      const double got = (acc[0][e] + acc[1][e]) + (acc[2][e] + acc[3][e]);
      double c4[4] = {0, 0, 0, 0};
      for(int k = 0; k < 64; ++k) c4[k & 3] = fma(X(g, k), Y(k, c), c4[k & 3]);
      d4 += got == (c4[0] + c4[1]) + (c4[2] + c4[3]);
      ++d4tot;
    }
  }
  atomicAdd(counts + 0, seq);
  atomicAdd(counts + 1, pair);
  atomicAdd(counts + 2, tot);
  atomicAdd(counts + 3, d4);
  atomicAdd(counts + 4, d4tot);
}

} // namespace

extern "C" int jrlqp_probe_dmma(int32_t device, int32_t reps, double * out6)
{
  if(!out6 || reps < 1) return JRLQP_ERR_ARG;
  if(cudaSetDevice(device) != cudaSuccess) return JRLQP_ERR_CUDA;
  cudaDeviceProp prop;
  if(cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) return JRLQP_ERR_CUDA;
  const int grid = prop.multiProcessorCount; // one 4-warp CTA per SM, as the n = 128 kernel runs
  const int smem = (PK * PM + PK * PN) * 8;
  double *c1 = nullptr, *c2 = nullptr;
  unsigned long long * cnt = nullptr;
  bool ok = cudaMalloc(&c1, sizeof(double) * grid * PM * PN) == cudaSuccess && cudaMalloc(&c2, sizeof(double) * grid * PM * PN) == cudaSuccess &&
            cudaMalloc(&cnt, 5 * sizeof(unsigned long long)) == cudaSuccess;
  ok = ok && cudaFuncSetAttribute(gemm_fp64_pipe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) == cudaSuccess;
  ok = ok && cudaFuncSetAttribute(gemm_dmma, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) == cudaSuccess;
  int rc = JRLQP_ERR_CUDA;
  if(ok)
  {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float ms[2] = {0, 0};
    for(int which = 0; which < 2; ++which)
    {
      for(int pass = 0; pass < 2; ++pass)
      {
        cudaEventRecord(e0);
        if(which == 0)
          gemm_fp64_pipe<<<grid, 128, smem>>>(c1, reps, 42ull);
        else
          gemm_dmma<<<grid, 128, smem>>>(c2, reps, 42ull);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms[which], e0, e1);
      }
    }
    const double flop = 2.0 * PM * PN * PK * (double)reps * grid;
    out6[0] = flop / (ms[0] * 1e-3) / 1e9;
    out6[1] = flop / (ms[1] * 1e-3) / 1e9;
    std::vector<double> h1((size_t)grid * PM * PN), h2(h1.size());
    cudaMemcpy(h1.data(), c1, sizeof(double) * h1.size(), cudaMemcpyDeviceToHost);
    cudaMemcpy(h2.data(), c2, sizeof(double) * h2.size(), cudaMemcpyDeviceToHost);
    size_t same = 0;
    for(size_t i = 0; i < h1.size(); ++i) same += h1[i] == h2[i];
    out6[5] = same == h1.size() ? 1.0 : (double)same / (double)h1.size();
    cudaMemset(cnt, 0, 5 * sizeof(unsigned long long));
    dmma_order<<<64, 32>>>(7ull, 256, cnt);
    unsigned long long hc[5];
    ok = cudaMemcpy(hc, cnt, sizeof(hc), cudaMemcpyDeviceToHost) == cudaSuccess;
    out6[2] = (double)hc[0] / (double)hc[2];
    out6[3] = (double)hc[1] / (double)hc[2];
    out6[4] = (double)hc[3] / (double)hc[4];
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    rc = ok && cudaGetLastError() == cudaSuccess ? JRLQP_OK : JRLQP_ERR_CUDA;
  }
  cudaFree(c1);
  cudaFree(c2);
  cudaFree(cnt);
  return rc;
}
