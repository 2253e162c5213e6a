/*
 * h9_pack.cu -- K4: forcing ingest.  The host hands over the seven PGF arrays
 * exactly as READ_PGF leaves them, (lon_c,lat_c,ndays) with x fastest and the
 * day slowest (READ_NET_CDF_3DR.f90:95-97).  This kernel does the land compaction and
 * the (x,y,day) -> [day][field][cell] re-layout the time-stepping kernel streams.  When
 * the host arrays are page-locked it reads them in place over PCIe (zero-copy gather:
 * only the land cells, about a quarter of the grid, cross the bus); pageable arrays are
 * first copied tile by tile through a pinned staging ring.
 */
#include <cuda_runtime.h>

#include <algorithm>

#include "h9_device.h"

namespace h9 {

struct PackSrc {
  const float* plane[kForcingPlanes]; /* each (ndays, ngrid): device memory or mapped pinned host memory */
};

__global__ void __launch_bounds__(256)
pack_forcing_kernel(const PackSrc src, int ndays, size_t ngrid,
                    const int32_t* __restrict__ cell_xy, int nc, int ncs,
                    float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int d = blockIdx.y;
  if (c >= ncs) return;
  /* padding cells replicate the last land cell so that staged tiles hold finite data */
  const int cell = cell_xy[c < nc ? c : nc - 1];
  float v[kForcingPlanes];
#pragma unroll
  for (int p = 0; p < kForcingPlanes; ++p) v[p] = __ldg(src.plane[p] + (size_t)d * ngrid + cell);
#pragma unroll
  for (int p = 0; p < kForcingPlanes; ++p) out[((size_t)d * kForcingPlanes + p) * ncs + c] = v[p];
}

int launch_pack_forcing(void* stream, const float* const planes[kForcingPlanes], int ndays,
                        size_t ngrid, const int32_t* d_cell_xy, int nc, int ncs, float* d_out) {
  if (nc <= 0 || ndays <= 0) return 0;
  PackSrc src;
  for (int p = 0; p < kForcingPlanes; ++p) src.plane[p] = planes[p];
  dim3 grid((ncs + 255) / 256, ndays);
  pack_forcing_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, ndays, ngrid, d_cell_xy, nc, ncs,
                                                              d_out);
  return (int)cudaGetLastError();
}

/* N4: INIT.f90:573-633.  One block per half-degree cell: the 60x60 fine cells of the four
 * fields are brought into shared memory with coalesced loads, then four threads (one per
 * field) add them up in the reference's order (x1 outer, y1 inner, only where theta_s_in >= 0)
 * so that the float sums are bit-identical to the Fortran loop. */
__global__ void __launch_bounds__(256)
regrid_soil_kernel(const float* __restrict__ in, int lon_c, int rows, float* __restrict__ out) {
  extern __shared__ float tile_raw[];
  float (*tile)[60 * 61] = reinterpret_cast<float (*)[60 * 61]>(tile_raw); /* [field][y1][x1], padded rows */
  const int x = blockIdx.x, y = blockIdx.y;
  const size_t fw = (size_t)lon_c * 60, plane = fw * (size_t)rows * 60;
  for (int k = threadIdx.x; k < 4 * 3600; k += blockDim.x) {
    const int f = k / 3600, r = k % 3600, y1 = r / 60, x1 = r % 60;
    tile[f][y1 * 61 + x1] = in[f * plane + ((size_t)y * 60 + y1) * fw + (size_t)x * 60 + x1];
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    const int f = threadIdx.x;
    float sum = 0.0f;
    int j = 0;
    for (int x1 = 0; x1 < 60; ++x1)
      for (int y1 = 0; y1 < 60; ++y1)
        if (tile[0][y1 * 61 + x1] >= 0.0f) { /* :584 the mask is theta_s_in for all four */
          sum = sum + tile[f][y1 * 61 + x1];
          ++j;
        }
    if (j > 0) sum = sum / (float)j; /* :593-598 */
    float v;
    if (f == 0) v = sum / 1.0E3f;               /* theta_s :613 */
    else if (f == 1) v = 10.0f * sum / 86400.0f; /* hksat   :614 */
    else if (f == 2) v = 1.0f / fmaxf(sum / 1.0E3f, 1.0E-8f); /* lambda :615,624 -> bsw :628 */
    else v = 10.0f * sum;                        /* psi_s   :616 */
    out[((size_t)f * rows + y) * lon_c + x] = v;
  }
}

int launch_regrid_soil(void* stream, const float* d_in, int lon_c, int rows, float* d_out) {
  if (lon_c <= 0 || rows <= 0) return 0;
  const size_t shm = 4 * 60 * 61 * sizeof(float); /* 58.6 KB: above the 48 KB static limit */
  cudaError_t e = cudaFuncSetAttribute(regrid_soil_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
  if (e != cudaSuccess) return (int)e;
  regrid_soil_kernel<<<dim3(lon_c, rows), 256, shm, (cudaStream_t)stream>>>(d_in, lon_c, rows, d_out);
  return (int)cudaGetLastError();
}

__global__ void clear_u32_kernel(uint32_t* p, size_t n, uint32_t v) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

int launch_clear_u32(void* stream, uint32_t* p, size_t n, uint32_t v) {
  if (n == 0) return 0;
  clear_u32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p, n, v);
  return (int)cudaGetLastError();
}

/* K5: per-year budget partial sums over the cells of this ctx, FP64 on purpose
 * (the summation order depends on the shard count).  out[0..7] = sum of soil
 * water (current h2osoi_liq, all layers), sum of wa, sum of annual rnf, sum of
 * annual npp, sum of annual mean plant mass, number of cells, sum of annual
 * mean total soil water, number of faulted cells.  Not in the reference
 * (SURVEY.md section 8e); feeds the NCCL all-reduce of the budget line. */
__global__ void __launch_bounds__(256)
budget_kernel(const DevArrays a, int iyr, double* __restrict__ out) {
  __shared__ double sh[8][8];
  double v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const float* ann = a.annual + (size_t)(iyr - 1) * kAnnualPlanes * a.ncs;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < a.nc; c += gridDim.x * blockDim.x) {
    double w = 0.0;
#pragma unroll
    for (int i = 0; i < NL; ++i) w += (double)a.h2o[(size_t)c * NL + i];
    v[0] += w;
    v[1] += (double)a.wa[c];
    v[2] += (double)ann[(size_t)2 * a.ncs + c];
    v[3] += (double)ann[c];
    v[4] += (double)ann[(size_t)1 * a.ncs + c];
    v[5] += 1.0;
    v[6] += (double)ann[(size_t)4 * a.ncs + c];
    v[7] += a.fault[c] ? 1.0 : 0.0;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    double x = v[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) sh[q][warp] = x;
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    double x = 0.0;
    for (int w = 0; w < 8; ++w) x += sh[threadIdx.x][w];
    atomicAdd(out + threadIdx.x, x);
  }
}

int launch_budget(void* stream, const DevArrays& a, int iyr, double* d_budget) {
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(d_budget, 0, 8 * sizeof(double), st);
  if (e != cudaSuccess) return (int)e;
  if (a.nc <= 0) return 0;
  const int blocks = std::min(148, (a.nc + 255) / 256);
  budget_kernel<<<blocks, 256, 0, st>>>(a, iyr, d_budget);
  return (int)cudaGetLastError();
}

} /* namespace h9 */
