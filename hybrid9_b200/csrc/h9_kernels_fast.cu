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
#include "h9_physics_fast_tp.cuh"

namespace h9 {

#ifdef H9_CYCLE_BUDGET
/* cycles per segment of cell 0's sub-steps, summed over the launch (tools/cycle_budget.py) */
__device__ unsigned long long g_cycle_budget[kTickSegs + 2];
/* per warp of the launch: total cycles of lane 0, and how often its sub-steps took the
 * cascade / dryness-repair branch */
__device__ unsigned g_warp_cycles[4096], g_warp_repairs[4096], g_warp_smid[4096], g_warp_general[4096],
    g_warp_slow[4096];
#endif

template <int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB)
days_kernel_fast(const DevArrays a, const __grid_constant__ Geo g, int ndays,
                 const int32_t* __restrict__ year_index, const ForcingView fv, int cur_year,
                 int nt, unsigned long long step0, int cells_per_block,
                 const int32_t* __restrict__ cell_index) {
  extern __shared__ float4 smem[];
  __shared__ float s_geo[kGeoDynFloats];
  GeoDyn::fill(s_geo, g); /* before any thread leaves: it ends with the kernel's only barrier */
  const GeoDyn gd{s_geo};
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
  const float evap_on = a.real_evap ? 1.0f : 0.0f;

  /* compact forcing [day][7][ncs], or the grid tile as the host holds it gathered through
   * the land index (no separate pack pass on the h9_run_days pipeline) */
  const size_t fcell = (size_t)(cell_index ? __ldg(cell_index + c) : c);
  Forcing fnext = load_forcing(fv, fcell);
  int iy_next = __ldg(year_index);

#ifdef H9_CYCLE_BUDGET
  Ticks tk;
#pragma unroll
  for (int k = 0; k < kTickSegs; ++k) tk.acc[k] = 0u;
  unsigned t_loop = 0u, t_all0 = 0u;
  asm volatile("mov.u32 %0, %%clock;" : "=r"(t_all0)::"memory");
#endif
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

    /* fault bookkeeping.  The build with all registers (small shards: one warp per scheduler,
     * nothing hides the wait for the step's last result) keeps the day's first fault in three
     * registers and tests once per day; the 128-register build tests every sub-step. */
    constexpr bool kDeferFault = (MINB == 1);
    uint32_t ft_day = 0u, ft_first = 0u;
    int ft_ns = 0;
    float ft_imb = 0.0f;
    for (int ns = 0; ns < g.nisurf; ++ns) { /* HYBRID9.f90:193-211 */
      StepOut so;
      /* the build with all registers picks, per warp and sub-step, the step without any
       * water-table-in-column code when every cell of the warp has a deep table (a uniform
       * branch), else the general straight-line step; the 128-register build keeps the step with
       * the fewest executed instructions (h9_physics_fast_tp.cuh) */
      uint32_t ft;
      if (MINB == 1) {
        const bool all_deep = __all_sync(__activemask(), !(s.zwt <= g.zim[NL]));
        ft = all_deep ? hydrology_step_fast<kStepAllDeep>(g, gd, tbl, day, s, so H9_TICKS_ARG)
                      : hydrology_step_fast<kStepGeneral>(g, gd, tbl, day, s, so H9_TICKS_ARG);
      } else {
        ft = hydrology_step_fast_tp(g, tbl, day, s, so);
      }
      if (kDeferFault) {
        const bool first = (ft != 0u) && (ft_day == 0u);
        ft_first = first ? ft : ft_first;
        ft_ns = first ? ns : ft_ns;
        ft_imb = first ? so.imbalance : ft_imb;
        ft_day |= ft;
      } else if (ft) {
        record_fault(a, c, sticky, ft, step0 + (unsigned long long)d * g.nisurf + ns, so.imbalance);
      }
      /* H9_OPT_REAL_EVAP: evap_sum stays 0 otherwise, as in the reference (HYBRID9.f90:137,276) */
      evap_sum = fmaf(evap_on, so.qflx_evap_grnd + so.qflx_tran_veg_col, evap_sum);
      H9_TICK(6); /* fault test, loop */
    }
    if (kDeferFault && ft_day) {
      record_fault(a, c, sticky, ft_first, step0 + (unsigned long long)d * g.nisurf + ft_ns, ft_imb);
      if ((sticky | ft_day) != sticky) { /* further bits raised later in the day */
        atomicOr(a.any_fault, ft_day);
        sticky |= ft_day;
      }
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
#ifdef H9_CYCLE_BUDGET
  unsigned warp_repairs = tk.acc[kTickSegs - 1]; /* most often any lane took the branch */
  unsigned warp_general = tk.acc[kTickSegs - 2];
  /* per lane: sub-steps that left the straight-line tail because a Drainage loop went on
   * (recharge: s1, baseflow: s2) or the bottom layer had to search upward (s3) */
  unsigned s1 = tk.acc[kTickSegs - 3], s2 = tk.acc[kTickSegs - 4], s3 = tk.acc[kTickSegs - 5];
  for (int o = 16; o > 0; o >>= 1) {
    warp_repairs = max(warp_repairs, __shfl_xor_sync(__activemask(), warp_repairs, o));
    s1 = max(s1, __shfl_xor_sync(__activemask(), s1, o));
    s2 = max(s2, __shfl_xor_sync(__activemask(), s2, o));
    s3 = max(s3, __shfl_xor_sync(__activemask(), s3, o));
  }
  if ((threadIdx.x & 31) == 0) {
    unsigned t_w;
    asm volatile("mov.u32 %0, %%clock;" : "=r"(t_w)::"memory");
    const int wg = (blockIdx.x * BLOCK + threadIdx.x) >> 5;
    if (wg < 4096) {
      g_warp_cycles[wg] = t_w - t_all0;
      g_warp_repairs[wg] = warp_repairs;
      g_warp_general[wg] = warp_general;
      g_warp_slow[wg] = s1;
      g_warp_slow[1024 + wg % 1024] = s2;
      g_warp_slow[2048 + wg % 1024] = s3;
      unsigned smid, warpid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      asm volatile("mov.u32 %0, %%warpid;" : "=r"(warpid));
      g_warp_smid[wg] = (smid << 8) | (warpid & 0xffu);
    }
  }
  if (c == a.budget_cell) {
    unsigned t_all1;
    asm volatile("mov.u32 %0, %%clock;" : "=r"(t_all1)::"memory");
    for (int k = 0; k < kTickSegs; ++k) g_cycle_budget[k] = tk.acc[k];
    g_cycle_budget[kTickSegs] = (unsigned long long)(t_all1 - t_all0);
    g_cycle_budget[kTickSegs + 1] = (unsigned long long)ndays * g.nisurf;
    (void)t_loop;
  }
#endif
}

__global__ void __launch_bounds__(128)
hydrology_step_kernel_fast(const DevArrays a, const __grid_constant__ Geo g, const ForcingView fv,
                           unsigned long long step0, const StepDiagArrays diag, int variant) {
  __shared__ float s_geo[kGeoDynFloats];
  GeoDyn::fill(s_geo, g);
  const GeoDyn gd{s_geo};
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
#ifdef H9_CYCLE_BUDGET
  Ticks tk;
#endif
  /* the step variants of the small-shard day kernel, chosen the same way (warp vote) */
  const bool all_deep = __all_sync(__activemask(), !(s.zwt <= g.zim[NL]));
  uint32_t ft;
  if (variant == kStepThroughput) {
    ft = hydrology_step_fast_tp(g, tbl, day, s, so);
  } else if (all_deep) {
    ft = hydrology_step_fast<kStepAllDeep>(g, gd, tbl, day, s, so H9_TICKS_ARG);
  } else {
    ft = hydrology_step_fast<kStepGeneral>(g, gd, tbl, day, s, so H9_TICKS_ARG);
  }
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

/* Launch shape of the thread-per-cell kernel.  The default (1064) adapts to the shard size: up
 * to 8 warps per SM the uncapped build (more registers, more instruction-level parallelism per
 * warp) is faster, above that the 128-register build keeps the whole shard resident in one wave.
 * A shard of at most 128 cells per SM goes out as ONE block of four warps per SM: the four land
 * on the four schedulers, whereas two 64-thread blocks on an SM can put two warps on one
 * scheduler and leave another idle (quarter of the 0.5 deg grid: 24.9 -> 21.7 ms per year).
 * Same PTX, same results for every shape. */
static void fast_shape(int nc, int block, int* bs, bool* capped) {
  int dev = 0, nsm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  const bool small_shard = (block == 1064) && (nc <= nsm * 8 * 32);
  *capped = block >= 1000 && !small_shard;
  *bs = block % 1000;
  if (small_shard && nc <= nsm * 128) *bs = 128;
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
  int bs = 64;
  bool capped = false;
  fast_shape(a.nc, block, &bs, &capped);
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
  if (block >= 2000) return "h9::days_kernel_fast<512,1> (balanced)";
  int bs = 64;
  bool capped = false;
  fast_shape(nc, block, &bs, &capped);
  if (bs == 32) return capped ? "h9::days_kernel_fast<32,16>" : "h9::days_kernel_fast<32,1>";
  if (bs == 128) return capped ? "h9::days_kernel_fast<128,4>" : "h9::days_kernel_fast<128,1>";
  return capped ? "h9::days_kernel_fast<64,8>" : "h9::days_kernel_fast<64,1>";
}

#ifdef H9_CYCLE_BUDGET
extern "C" int h9_debug_cycle_budget(unsigned long long* out) {
  return (int)cudaMemcpyFromSymbol(out, g_cycle_budget, sizeof(unsigned long long) * (kTickSegs + 2));
}
extern "C" int h9_debug_warp_cycles(unsigned* cycles, unsigned* repairs) {
  cudaError_t e = cudaMemcpyFromSymbol(cycles, g_warp_cycles, sizeof(unsigned) * 4096);
  if (e != cudaSuccess) return (int)e;
  return (int)cudaMemcpyFromSymbol(repairs, g_warp_repairs, sizeof(unsigned) * 4096);
}
extern "C" int h9_debug_warp_general(unsigned* general, unsigned* slow) {
  cudaError_t e = cudaMemcpyFromSymbol(general, g_warp_general, sizeof(unsigned) * 4096);
  if (e != cudaSuccess) return (int)e;
  return (int)cudaMemcpyFromSymbol(slow, g_warp_slow, sizeof(unsigned) * 4096);
}
extern "C" int h9_debug_warp_smid(unsigned* smid) {
  return (int)cudaMemcpyFromSymbol(smid, g_warp_smid, sizeof(unsigned) * 4096);
}
#endif

int launch_hydrology_step_fast(void* stream, const DevArrays& a, const Geo& g,
                               const ForcingView& fv, unsigned long long step0,
                               const StepDiagArrays& diag) {
  return launch_hydrology_step_fast_variant(stream, a, g, fv, step0, diag, 1064);
}

/* `block` as for the day kernel: the sub-step variant that kernel would run for this shard */
int launch_hydrology_step_fast_variant(void* stream, const DevArrays& a, const Geo& g,
                                       const ForcingView& fv, unsigned long long step0,
                                       const StepDiagArrays& diag, int block) {
  if (a.nc <= 0) return 0;
  int dev = 0, nsm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  const bool small_shard = (block == 1064) && (a.nc <= nsm * 8 * 32);
  const bool capped = block >= 1000 && block < 2000 && !small_shard;
  hydrology_step_kernel_fast<<<(a.nc + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      a, g, fv, step0, diag, capped ? (int)kStepThroughput : (int)kStepGeneral);
  return (int)cudaGetLastError();
}

int launch_grow_day_fast(void* stream, const DevArrays& a, const Geo& g, const float* d_tas,
                         const GrowDiagArrays& diag) {
  return launch_grow_day_t<MathFast>(stream, a, g, d_tas, diag);
}

} /* namespace h9 */
