// Latency / partial-warp microbenchmark for the lone-warp regime of the fast step (sm_100a).
//   1. MUFU.EX2 throughput of ONE warp per scheduler with 32 / 16 / 8 / 4 active lanes
//      (does the 4-lane XU skip inactive quarter-warps?)
//   2. packed FP32 (fma.rn.f32x2 / mul / add): issue cost and dependent latency
//   3. dependent-chain latencies: FFMA, FMNMX, MUFU ex2 / lg2 / rcp, SHFL, LDS, pow chain
// One block per SM, 128 threads (one warp per scheduler) unless stated.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void ex2(float& x) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x)); }
__device__ __forceinline__ void lg2(float& x) { asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(x)); }
__device__ __forceinline__ void rcp(float& x) { asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x)); }
__device__ __forceinline__ void ffma(float& x, float b, float c) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x) : "f"(b), "f"(c)); }
__device__ __forceinline__ void fmnmx(float& x, float c) { asm volatile("max.f32 %0, %0, %1;" : "+f"(x) : "f"(c)); }
__device__ __forceinline__ void ffma2(unsigned long long& x, unsigned long long b, unsigned long long c) {
  asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x) : "l"(b), "l"(c));
}
__device__ __forceinline__ void fmul2(unsigned long long& x, unsigned long long b) {
  asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(x) : "l"(b));
}
__device__ __forceinline__ void fadd2(unsigned long long& x, unsigned long long b) {
  asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(x) : "l"(b));
}
__device__ __forceinline__ unsigned long long pack(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ float lo_of(unsigned long long v) {
  float lo, hi;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
  return lo + hi;
}

// MODE 0: MUFU.EX2, ILP 8, lanes < nactive only
// MODE 1: FFMA2 ILP 8      MODE 2: FFMA2 dependent chain (ILP 1)
// MODE 3: FFMA chain       MODE 4: FMNMX chain     MODE 5: ex2 chain   MODE 6: lg2 chain  MODE 7: rcp chain
// MODE 8: SHFL chain       MODE 9: LDS chain       MODE 10: pow chain (max, lg2, mul, ex2)
// MODE 11: FMUL2 ILP 8     MODE 12: FADD2 ILP 8    MODE 13: FFMA ILP 8 (reference)
// MODE 14: ex2 ILP8 with every second lane active (stride-2 mask)
// MODE 15: FFMA ILP 2 ; MODE 16: FFMA2 ILP 2 ; MODE 17: ex2 ILP 2 ; MODE 18: ex2 ILP 4
template <int MODE>
__global__ void k(float* out, int iters, float b, float c, int nactive, int* perm) {
  __shared__ int sm[128];
  sm[threadIdx.x] = perm ? perm[threadIdx.x] : threadIdx.x;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  float a[8];
  unsigned long long p[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = 1.0f + 0.001f * (threadIdx.x + i);
    p[i] = pack(a[i], a[i] + 0.5f);
  }
  const unsigned long long pb = pack(b, b), pc = pack(c, c);
  int idx = threadIdx.x;
  const bool act = (MODE == 14) ? ((lane & 1) == 0 && (lane >> 1) < nactive) : (lane < nactive);
  if (act) {
    for (int it = 0; it < iters; ++it) {
      if (MODE == 0 || MODE == 14) {
#pragma unroll
        for (int i = 0; i < 8; ++i) ex2(a[i]);
      } else if (MODE == 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) ffma2(p[i], pb, pc);
      } else if (MODE == 2) {
#pragma unroll
        for (int i = 0; i < 8; ++i) ffma2(p[0], pb, pc);
      } else if (MODE == 3) {
#pragma unroll
        for (int i = 0; i < 8; ++i) ffma(a[0], b, c);
      } else if (MODE == 4) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { fmnmx(a[0], c); ffma(a[0], b, c); }
      } else if (MODE == 5) {
#pragma unroll
        for (int i = 0; i < 8; ++i) ex2(a[0]);
      } else if (MODE == 6) {
#pragma unroll
        for (int i = 0; i < 8; ++i) lg2(a[0]);
      } else if (MODE == 7) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { rcp(a[0]); ffma(a[0], b, c); }
      } else if (MODE == 8) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[0] = __shfl_xor_sync(0xffffffffu, a[0], 1);
      } else if (MODE == 9) {
#pragma unroll
        for (int i = 0; i < 8; ++i) idx = ((volatile int*)sm)[idx];
      } else if (MODE == 10) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { fmnmx(a[0], c); lg2(a[0]); ffma(a[0], b, c); ex2(a[0]); }
      } else if (MODE == 11) {
#pragma unroll
        for (int i = 0; i < 8; ++i) fmul2(p[i], pb);
      } else if (MODE == 12) {
#pragma unroll
        for (int i = 0; i < 8; ++i) fadd2(p[i], pc);
      } else if (MODE == 13) {
#pragma unroll
        for (int i = 0; i < 8; ++i) ffma(a[i], b, c);
      } else if (MODE == 15) {
#pragma unroll
        for (int i = 0; i < 8; ++i) ffma(a[i & 1], b, c);
      } else if (MODE == 16) {
#pragma unroll
        for (int i = 0; i < 8; ++i) ffma2(p[i & 1], pb, pc);
      } else if (MODE == 17) {
#pragma unroll
        for (int i = 0; i < 8; ++i) ex2(a[i & 1]);
      } else if (MODE == 18) {
#pragma unroll
        for (int i = 0; i < 8; ++i) ex2(a[i & 3]);
      }
    }
  }
  float s = (float)idx;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + lo_of(p[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static double g_ghz;
static int g_nsm;
static float* g_d;

template <int MODE>
void run(const char* name, int ops_per_iter, int nactive = 32, int threads = 128) {
  const int iters = 20000;
  k<MODE><<<g_nsm, threads>>>(g_d, 100, 1.0000001f, 1e-9f, nactive, nullptr);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<g_nsm, threads>>>(g_d, iters, 1.0000001f, 1e-9f, nactive, nullptr);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double cycles = ms * 1e-3 * g_ghz * 1e9;
  printf("%-44s lanes %2d warps/SMSP %d : %.2f cycles per warp-instruction\n", name, nactive, threads / 128,
         cycles / ((double)iters * ops_per_iter * (threads / 128)));
}

int main() {
  int dev = 0, khz = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&g_nsm, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  g_ghz = khz * 1e-6;
  printf("SMs %d, clock %.3f GHz (nominal; cycles assume it)\n", g_nsm, g_ghz);
  cudaMalloc(&g_d, sizeof(float) * g_nsm * 1024);
  for (int n : {32, 24, 16, 8, 4, 1}) run<0>("MUFU.EX2 ILP8, lanes [0,n)", 8, n);
  for (int n : {16, 8, 4}) run<14>("MUFU.EX2 ILP8, n even lanes", 8, n);
  run<0>("MUFU.EX2 ILP8, lanes [0,n), 2 warps/SMSP", 8, 16, 256);
  run<0>("MUFU.EX2 ILP8, lanes [0,n), 4 warps/SMSP", 8, 8, 512);
  run<13>("FFMA ILP8", 8);
  run<1>("FFMA2 (f32x2) ILP8", 8);
  run<11>("FMUL2 ILP8", 8);
  run<12>("FADD2 ILP8", 8);
  run<1>("FFMA2 ILP8 2 warps/SMSP", 8, 32, 256);
  run<1>("FFMA2 ILP8 4 warps/SMSP", 8, 32, 512);
  run<15>("FFMA ILP2", 8);
  run<16>("FFMA2 ILP2", 8);
  run<17>("MUFU.EX2 ILP2", 8);
  run<18>("MUFU.EX2 ILP4", 8);
  run<3>("FFMA dependent chain (latency)", 8);
  run<2>("FFMA2 dependent chain (latency)", 8);
  run<4>("FMNMX+FFMA dependent chain (per pair)", 8);
  run<5>("MUFU.EX2 dependent chain (latency)", 8);
  run<6>("MUFU.LG2 dependent chain (latency)", 8);
  run<7>("MUFU.RCP+FFMA dependent chain (per pair)", 8);
  run<8>("SHFL.BFLY dependent chain (latency)", 8);
  run<9>("LDS dependent chain (latency)", 8);
  run<10>("pow chain max,lg2,fma,ex2 (per pow)", 8);
  return 0;
}
