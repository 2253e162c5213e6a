// Issue-cost microbenchmark for the fast step's instruction mix (sm_100a).
// Each thread keeps 8 independent accumulators; one "group" = one op per accumulator (ILP 8).
// Reports cycles per warp-instruction per SMSP at 1/2/4 warps per scheduler.
#include <cstdio>
#include <cuda_runtime.h>

#define OP_FFMA3(x, y, z) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x) : "f"(y), "f"(z))
#define OP_FFMAI(x) asm volatile("fma.rn.f32 %0, %0, 0f3F800001, 0f3A83126F;" : "+f"(x))
#define OP_FMUL(x, y) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(x) : "f"(y))
#define OP_FADD(x, y) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x) : "f"(y))
#define OP_FMNMX(x, y) asm volatile("max.f32 %0, %0, %1;" : "+f"(x) : "f"(y))
#define OP_EX2(x) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x))
#define OP_LG2(x) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(x))
#define OP_RCP(x) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x))

template <int MODE>
__global__ void k(float* out, int iters, float b, float c) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = 1.0f + 0.001f * (threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) OP_FFMA3(a[i], b, c);
      if (MODE == 1) OP_FFMAI(a[i]);
      if (MODE == 2) OP_FMUL(a[i], b);
      if (MODE == 3) OP_FADD(a[i], c);
      if (MODE == 4) OP_FMNMX(a[i], c);
      if (MODE == 5) OP_EX2(a[i]);
      if (MODE == 6) { OP_FMUL(a[i], b); OP_FADD(a[i], c); }                       // 2 fma-pipe ops
      if (MODE == 7) { OP_FMUL(a[i], b); OP_FMNMX(a[i], c); }                      // fma + alu
      if (MODE == 8) { OP_FMUL(a[i], b); OP_FADD(a[i], c); OP_FMNMX(a[i], c); OP_FFMA3(a[i], b, c); } // 4 ops
      if (MODE == 9) { OP_FMUL(a[i], b); OP_FADD(a[i], c); OP_FMNMX(a[i], c); OP_FFMA3(a[i], b, c);
                       OP_FMUL(a[i], b); OP_FADD(a[i], c); OP_FMNMX(a[i], c); if (i == 0) OP_EX2(a[i]); } // 57 : 1
      if (MODE == 10) { OP_FMUL(a[i], b); OP_FADD(a[i], c); OP_FMNMX(a[i], c); OP_FFMA3(a[i], b, c);
                        OP_FMUL(a[i], b); OP_FADD(a[i], c); OP_FMNMX(a[i], c); if (i < 4) OP_EX2(a[i]); } // 56 : 4 (14:1)
      if (MODE == 11) { OP_FMUL(a[i], b); OP_FADD(a[i], c); OP_FMNMX(a[i], c); OP_FFMA3(a[i], b, c);
                        OP_FMUL(a[i], b); OP_FADD(a[i], c); OP_FMNMX(a[i], c); OP_EX2(a[i]); }            // 7 : 1
      if (MODE == 12) { OP_FMUL(a[i], b); OP_FADD(a[i], c); OP_FMNMX(a[i], c); OP_EX2(a[i]); }            // 3 : 1
      if (MODE == 13) { /* one pow per accumulator: clamp, lg2, *b, ex2, then 3 dependent FMA-pipe ops (6 : 2) */
        OP_FMNMX(a[i], c); OP_FADD(a[i], b); OP_LG2(a[i]); OP_FMUL(a[i], b); OP_EX2(a[i]); OP_FMUL(a[i], b); OP_FADD(a[i], c); OP_FMUL(a[i], b); }
      if (MODE == 14) { /* the same with 12 FMA/ALU ops per pow (12 : 2), the fast step's overall ratio is ~8.7 : 1 */
        OP_FMNMX(a[i], c); OP_FADD(a[i], b); OP_LG2(a[i]); OP_FMUL(a[i], b); OP_EX2(a[i]); OP_FMUL(a[i], b); OP_FADD(a[i], c); OP_FMUL(a[i], b);
        OP_FADD(a[i], c); OP_FMUL(a[i], b); OP_FMNMX(a[i], c); OP_FADD(a[i], c); OP_FMUL(a[i], b); OP_FADD(a[i], c); }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static const int kOps[15] = {8, 8, 8, 8, 8, 8, 16, 16, 32, 57, 60, 64, 32, 64, 112};
static const char* kName[15] = {"FFMA 3-reg", "FFMA imm", "FMUL", "FADD", "FMNMX", "MUFU.EX2", "FMUL+FADD",
                                "FMUL+FMNMX", "FMUL+FADD+FMNMX+FFMA", "mix 56 : 1 MUFU", "mix 56 : 4 MUFU",
                                "mix 7 : 1 MUFU", "mix 3 : 1 MUFU", "8 pows, 6 FMA : 2 MUFU", "8 pows, 12 FMA : 2 MUFU"};

template <int MODE>
void run(float* d, int nsm, double ghz) {
  for (int wps = 1; wps <= 4; wps *= 2) {
    const int threads = 128 * wps, iters = 20000;
    k<MODE><<<nsm, threads>>>(d, 100, 1.0000001f, 1e-9f);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<nsm, threads>>>(d, iters, 1.0000001f, 1e-9f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double cycles = ms * 1e-3 * ghz * 1e9;
    const double winstr_per_smsp = (double)iters * kOps[MODE] * wps;
    printf("%-24s warps/SMSP %d  cycles per warp-instr per SMSP %.3f\n", kName[MODE], wps, cycles / winstr_per_smsp);
  }
}

int main() {
  int dev = 0, nsm = 0, khz = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  const double ghz = khz * 1e-6;
  printf("SMs %d, clock %.3f GHz (nominal; cycles assume it)\n", nsm, ghz);
  float* d;
  cudaMalloc(&d, sizeof(float) * nsm * 512);
  run<0>(d, nsm, ghz); run<1>(d, nsm, ghz); run<2>(d, nsm, ghz); run<3>(d, nsm, ghz); run<4>(d, nsm, ghz);
  run<5>(d, nsm, ghz); run<6>(d, nsm, ghz); run<7>(d, nsm, ghz); run<8>(d, nsm, ghz); run<9>(d, nsm, ghz);
  run<10>(d, nsm, ghz); run<11>(d, nsm, ghz); run<12>(d, nsm, ghz); run<13>(d, nsm, ghz); run<14>(d, nsm, ghz);
  return 0;
}
