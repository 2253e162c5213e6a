/*
 * h9_kernels_fast.cu -- H9_MATH_FAST kernels.
 *
 * K3 days_kernel_fast: one thread per land cell; the cell's soil water (h2o, smp)
 * and scalars stay in registers for the whole batch of days, its state-independent
 * per-layer constants and root fractions sit in a shared-memory column
 * (h9_physics_fast.cuh), the next day's forcing is prefetched behind the current
 * day's NISURF sub-steps, and the annual accumulators are touched once per day.
 * Compiled for <=128 registers per thread (16 resident warps per SM) by default.
 * K1/K2 are the one-routine entries used by the parity tests.
 */
#include <cstdlib>
#include "h9_kernels.cuh"
#include "h9_physics_fast.cuh"

namespace h9 {

template <int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB)
days_kernel_fast(const DevArrays a, const __grid_constant__ Geo g, int ndays,
                 const int32_t* __restrict__ year_index, const ForcingView fv, int cur_year,
                 int nt, unsigned long long step0, int cells_per_block,
                 const int32_t* __restrict__ cell_index) {
  extern __shared__ float4 smem[];
  /* cells_per_block <= BLOCK: the balanced launch gives every SM the same number of cells.
   * Negative: -lanes, partial warps for small shards -- only the first `lanes` lanes of each
   * warp hold a cell, so that a shard of a few thousand cells still puts one warp on every
   * scheduler (the MUFU unit and the issue slot are per scheduler, not per lane) */
  int c;
  if (cells_per_block < 0) {
    const int lanes = -cells_per_block, lane = threadIdx.x & 31;
    c = (blockIdx.x * (BLOCK / 32) + (threadIdx.x >> 5)) * lanes + lane;
    if (lane >= lanes) return;
  } else {
    c = blockIdx.x * cells_per_block + threadIdx.x;
    if (threadIdx.x >= cells_per_block) return;
  }
  if (c >= a.nc) return; /* no barriers below */
  const CellTable<BLOCK> tbl{smem + threadIdx.x};

  FastState s;
  State gs; /* GROW's view of the cell (once per day) */
  {
    Params p;
    load_params(a, c, p);
    load_state(a, c, gs);
    tbl.init(g, p, gs.rootr);
#pragma unroll
    for (int i = 0; i < NL; ++i) {
      s.h2o[i] = gs.h2o[i];
      s.smp[i] = gs.smp[i];
    }
    s.zwt = gs.zwt;
    s.wa = gs.wa;
    s.rnf_sum = gs.rnf_sum;
  }
  uint32_t sticky = a.fault[c];
  float evap_sum = a.real_evap ? a.evap_sum[c] : 0.0f;

  /* compact forcing [day][7][ncs], or the grid tile as the host holds it gathered through
   * the land index (no separate pack pass on the h9_run_days pipeline) */
  const size_t fcell = (size_t)(cell_index ? __ldg(cell_index + c) : c);
  Forcing fnext = load_forcing(fv, fcell);
  int iy_next = __ldg(year_index);

  for (int d = 0; d < ndays; ++d) {
    const Forcing f = fnext;
    const int iy = iy_next;
    if (d + 1 < ndays) { /* prefetch the next day's forcing behind this day's sub-steps */
      fnext = load_forcing(fv, fcell + (size_t)(d + 1) * fv.day_stride);
      iy_next = __ldg(year_index + d + 1);
    }
    const bool new_year = (iy != cur_year); /* HYBRID9.f90:134-146 */
    if (new_year) {
      cur_year = iy;
      nt = 0;
      s.rnf_sum = 0.0f;
      evap_sum = 0.0f;
    }

    DayFast day;
    float tas;
    day_setup_fast(g, f, gs.lai, gs.lai_litter, day, tas);

    for (int ns = 0; ns < g.nisurf; ++ns) { /* HYBRID9.f90:193-211 */
      StepOut so;
      const uint32_t ft = hydrology_step_fast(g, tbl, day, s, so);
      if (ft)
        record_fault(a, c, sticky, ft, step0 + (unsigned long long)d * g.nisurf + ns, so.imbalance);
      if (a.real_evap) evap_sum += so.qflx_evap_grnd + so.qflx_tran_veg_col; /* H9_OPT_REAL_EVAP */
    }

    /* GROW (HYBRID9.f90:217): reads smp and rootr, rewrites rootr */
#pragma unroll
    for (int i = 0; i < NL; ++i) {
      gs.smp[i] = s.smp[i];
      gs.rootr[i] = tbl.rootr(i);
    }
    GrowOut go;
    grow_day<MathFast>(g, tas, gs, go);
#pragma unroll
    for (int i = 0; i < NL; ++i) tbl.set_rootr(i, gs.rootr[i]);

    /* daily accumulators HYBRID9.f90:242-253: read-modify-write in L2-resident global memory */
    float npp_sum = new_year ? 0.0f : a.npp_sum[c];
    float plant_mass_sum = new_year ? 0.0f : a.plant_mass_sum[c];
    float h2osoi_sum_total = new_year ? 0.0f : a.h2osoi_sum_total[c];
    float theta_sum[NL];
    if (new_year) {
#pragma unroll
      for (int i = 0; i < NL; ++i) theta_sum[i] = 0.0f;
    } else {
      load8(a.theta_sum, c, theta_sum);
    }
    if (gs.nplants >= 1) plant_mass_sum += gs.plant_mass;
    npp_sum += go.npp;
#pragma unroll
    for (int i = 0; i < NL; ++i) {
      theta_sum[i] += fmaxf(s.h2o[i], 1.0E-6f) * g.rdzw[i + 1];
      h2osoi_sum_total += s.h2o[i];
    }
    nt += 1;
    a.npp_sum[c] = npp_sum;
    a.plant_mass_sum[c] = plant_mass_sum;
    a.h2osoi_sum_total[c] = h2osoi_sum_total;
    store8(a.theta_sum, c, theta_sum);

    /* year-end means HYBRID9.f90:263-291 (also at the end of a batch) */
    const bool close = (d + 1 == ndays) || (iy_next != iy);
    if (close && iy >= 1 && iy <= a.nyr) {
      float* out = a.annual + ((size_t)(iy - 1) * kAnnualPlanes) * a.ncs + c;
      const float rnt = MathFast::rcp((float)nt);
      out[0] = npp_sum;
      out[(size_t)1 * a.ncs] = plant_mass_sum * rnt;
      out[(size_t)2 * a.ncs] = s.rnf_sum * MathFast::rcp((float)(nt * g.nisurf));
      /* the reference's evap_sum never accumulates: axy_evap == 0 unless H9_OPT_REAL_EVAP */
      out[(size_t)3 * a.ncs] = a.real_evap ? evap_sum * MathFast::rcp((float)(nt * g.nisurf)) : 0.0f;
      out[(size_t)4 * a.ncs] = h2osoi_sum_total * rnt;
#pragma unroll
      for (int i = 0; i < NL; ++i) out[(size_t)(5 + i) * a.ncs] = theta_sum[i] * rnt;
    }
  }

#pragma unroll
  for (int i = 0; i < NL; ++i) {
    gs.h2o[i] = s.h2o[i];
    gs.smp[i] = s.smp[i];
  }
  gs.zwt = s.zwt;
  gs.wa = s.wa;
  gs.rnf_sum = s.rnf_sum;
  store_hydro_state(a, c, gs);
  store_grow_state(a, c, gs);
  a.fault[c] = sticky;
  if (a.real_evap) a.evap_sum[c] = evap_sum;
}

__global__ void __launch_bounds__(128)
hydrology_step_kernel_fast(const DevArrays a, const __grid_constant__ Geo g, const ForcingView fv,
                           unsigned long long step0, const StepDiagArrays diag) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.nc) return;
  float4 col[kFastFloatsPerCell / 4];
  const CellTable<1> tbl{col};
  Params p;
  State gs;
  load_params(a, c, p);
  load_state(a, c, gs);
  tbl.init(g, p, gs.rootr);
  FastState s;
#pragma unroll
  for (int i = 0; i < NL; ++i) {
    s.h2o[i] = gs.h2o[i];
    s.smp[i] = gs.smp[i];
  }
  s.zwt = gs.zwt;
  s.wa = gs.wa;
  s.rnf_sum = gs.rnf_sum;
  uint32_t sticky = a.fault[c];
  const Forcing f = load_forcing(fv, (size_t)c);
  DayFast day;
  float tas;
  day_setup_fast(g, f, gs.lai, gs.lai_litter, day, tas);
  StepOut so;
  const uint32_t ft = hydrology_step_fast(g, tbl, day, s, so);
  if (ft) record_fault(a, c, sticky, ft, step0, so.imbalance);
#pragma unroll
  for (int i = 0; i < NL; ++i) {
    gs.h2o[i] = s.h2o[i];
    gs.smp[i] = s.smp[i];
  }
  gs.zwt = s.zwt;
  gs.wa = s.wa;
  gs.rnf_sum = s.rnf_sum;
  store_hydro_state(a, c, gs);
  a.fault[c] = sticky;
  if (diag.theta) {
    float th[NL];
#pragma unroll
    for (int i = 0; i < NL; ++i) th[i] = fmaxf(s.h2o[i], 1.0E-6f) * g.rdzw[i + 1];
    store8(diag.theta, c, th);
  }
  if (diag.qflx_tran_veg_col) diag.qflx_tran_veg_col[c] = so.qflx_tran_veg_col;
  if (diag.qflx_evap_grnd) diag.qflx_evap_grnd[c] = so.qflx_evap_grnd;
  if (diag.rnf_inc) diag.rnf_inc[c] = so.rnf_inc;
  if (diag.w_imbalance) diag.w_imbalance[c] = so.imbalance;
  if (diag.jwt) diag.jwt[c] = so.jwt;
}

int launch_days_fast(void* stream, const DevArrays& a, const Geo& g, int ndays,
                     const int32_t* d_year_index, const ForcingView& fv, int cur_year, int nt,
                     unsigned long long step0, int block, const int32_t* d_cell_index) {
  cudaStream_t st = (cudaStream_t)stream;
  if (a.nc <= 0 || ndays <= 0) return 0;
  /* block: threads per block (32/64/128), +1000 for the <=128-register variant;
   * 2000 = balanced: one wide block per SM and wave, every SM gets the same cell count */
  if (block >= 2000) {
    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    const int waves = (a.nc + nsm * 512 - 1) / (nsm * 512);
    const int nblk = nsm * waves;
    const int cpb = (a.nc + nblk - 1) / nblk;
    const int threads = (cpb + 31) / 32 * 32;
    const size_t shm = (size_t)kFastFloatsPerCell * 512 * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(days_kernel_fast<512, 1>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
    if (e != cudaSuccess) return (int)e;
    days_kernel_fast<512, 1><<<(a.nc + cpb - 1) / cpb, threads, shm, st>>>(
        a, g, ndays, d_year_index, fv, cur_year, nt, step0, cpb, d_cell_index);
    return (int)cudaGetLastError();
  }
  /* the default (1064) adapts to the shard size: up to 8 warps per SM the uncapped variant
   * (more registers, more instruction-level parallelism per warp) is faster; above that the
   * 128-register variant keeps the whole shard resident in one wave.  Same PTX, same results. */
  int dev = 0, nsm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  const bool small_shard = (block == 1064) && (a.nc <= nsm * 8 * 32);
  const bool capped = block >= 1000 && !small_shard;
  const int bs = block % 1000;
  int lanes = 32;
  if (const char* e = getenv("H9_LANES")) lanes = atoi(e) > 0 && atoi(e) <= 32 ? atoi(e) : 32;
#define H9_LAUNCH(BS, MINB)                                                                    \
  do {                                                                                         \
    const size_t shm = (size_t)kFastFloatsPerCell * BS * sizeof(float);                        \
    cudaError_t e = cudaFuncSetAttribute(days_kernel_fast<BS, MINB>,                           \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm); \
    if (e != cudaSuccess) return (int)e;                                                       \
    const int cpb = (BS / 32) * lanes;                                                         \
    days_kernel_fast<BS, MINB><<<(a.nc + cpb - 1) / cpb, BS, shm, st>>>(                       \
        a, g, ndays, d_year_index, fv, cur_year, nt, step0, lanes == 32 ? BS : -lanes,         \
        d_cell_index);                                                                         \
  } while (0)
  if (bs == 32) {
    if (capped) H9_LAUNCH(32, 16); else H9_LAUNCH(32, 1);
  } else if (bs == 128) {
    if (capped) H9_LAUNCH(128, 4); else H9_LAUNCH(128, 1);
  } else {
    if (capped) H9_LAUNCH(64, 8); else H9_LAUNCH(64, 1);
  }
#undef H9_LAUNCH
  return (int)cudaGetLastError();
}

const char* days_variant_fast(int nc, int block) {
  int dev = 0, nsm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  if (block >= 2000) return "h9::days_kernel_fast<512,1> (balanced)";
  const bool small_shard = (block == 1064) && (nc <= nsm * 8 * 32);
  const bool capped = block >= 1000 && !small_shard;
  const int bs = block % 1000;
  if (bs == 32) return capped ? "h9::days_kernel_fast<32,16>" : "h9::days_kernel_fast<32,1>";
  if (bs == 128) return capped ? "h9::days_kernel_fast<128,4>" : "h9::days_kernel_fast<128,1>";
  return capped ? "h9::days_kernel_fast<64,8>" : "h9::days_kernel_fast<64,1>";
}

int launch_hydrology_step_fast(void* stream, const DevArrays& a, const Geo& g,
                               const ForcingView& fv, unsigned long long step0,
                               const StepDiagArrays& diag) {
  if (a.nc <= 0) return 0;
  hydrology_step_kernel_fast<<<(a.nc + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      a, g, fv, step0, diag);
  return (int)cudaGetLastError();
}

int launch_grow_day_fast(void* stream, const DevArrays& a, const Geo& g, const float* d_tas,
                         const GrowDiagArrays& diag) {
  return launch_grow_day_t<MathFast>(stream, a, g, d_tas, diag);
}

} /* namespace h9 */
