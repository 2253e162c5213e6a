/*
 * h9_kernels_pair.cu -- H9_MATH_FAST kernels with two lanes per land cell (h9_physics_pair.cuh),
 * for shards of up to ~9.5k cells (16 cells per warp on 148 x 4 schedulers): one latitude band
 * of the 0.5 deg grid on 8 GPUs, the regional block.
 *
 * K3p days_kernel_pair: the loop nest HYBRID9.f90:120-295 like days_kernel_fast; the cell's
 * soil column is split over the lane pair (layers 1..4 / 8..5 + aquifer), per-cell scalars and
 * the daily bookkeeping (GROW, accumulators, annual means) are done on the gathered column,
 * stored by the even lane.  No lane ever exits early: the shuffles use the full mask.
 * K1p hydrology_step_kernel_pair: one HYDROLOGY call, the 1:1 parity target of the pair
 * sub-step.
 */
#include <cstdlib>

#include "h9_kernels.cuh"
#include "h9_physics_pair.cuh"

namespace h9 {

#ifdef H9_CYCLE_BUDGET
/* per warp of the launch (tools/cycle_budget.py --pair): cycles in block A / tail of the
 * all-deep and of the general step, sub-steps on each, total */
__device__ unsigned g_pair_warp[8][4096];
extern "C" int h9_debug_pair_budget(unsigned* out) {
  return (int)cudaMemcpyFromSymbol(out, g_pair_warp, sizeof(unsigned) * 8 * 4096);
}
#endif

namespace {

/* the lane's half of a [ncs][8] field: layers 1..4 in order (h = 0), layers 8..5 (h = 1) */
__device__ __forceinline__ void load_half(const float* base, int c, int h, float (&v)[NH]) {
  const float4 q = reinterpret_cast<const float4*>(base)[2 * (size_t)c + h];
  v[0] = h ? q.w : q.x;
  v[1] = h ? q.z : q.y;
  v[2] = h ? q.y : q.z;
  v[3] = h ? q.x : q.w;
}

__device__ __forceinline__ void store_half(float* base, int c, int h, const float (&v)[NH]) {
  reinterpret_cast<float4*>(base)[2 * (size_t)c + h] =
      h ? make_float4(v[3], v[2], v[1], v[0]) : make_float4(v[0], v[1], v[2], v[3]);
}

} /* namespace */

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK, 1)
days_kernel_pair(const DevArrays a, const __grid_constant__ Geo g, int ndays,
                 const int32_t* __restrict__ year_index, const ForcingView fv, int cur_year,
                 int nt, unsigned long long step0, const int32_t* __restrict__ cell_index) {
  extern __shared__ float4 smem[];
  __shared__ float s_geo[kGeoDynFloats];
  GeoDyn::fill(s_geo, g);
  const GeoDyn gd{s_geo};
  const int lane = threadIdx.x & 31;
  const int h = threadIdx.x & 1;
  const int cell = (blockIdx.x * BLOCK + threadIdx.x) >> 1;
  const bool valid = cell < a.nc;
  const int c = valid ? cell : a.nc - 1; /* surplus lanes shadow the last cell and store nothing */
  const bool writer = valid && h == 0;
  const unsigned pmask = 3u << (lane & ~1);
  const PairTable<BLOCK> tbl{smem + threadIdx.x};
  PairGeo pg;
  pg.init(g, h);

  PairState s;
  State gs; /* GROW's view of the cell (once per day): whole column, both lanes */
  {
    Params p;
    load_params(a, c, p);
    load_state(a, c, gs);
    tbl.init(g, p, gs.rootr, h);
    load_half(a.h2o, c, h, s.h2o);
    load_half(a.smp, c, h, s.smp);
    s.zwt = gs.zwt;
    s.wa = gs.wa;
    s.rnf_sum = gs.rnf_sum;
  }
  uint32_t sticky = a.fault[c];
  float evap_sum = a.real_evap ? a.evap_sum[c] : 0.0f;
  const float evap_on = a.real_evap ? 1.0f : 0.0f;

  const size_t fcell = (size_t)(cell_index ? __ldg(cell_index + c) : c);
  Forcing fnext = load_forcing(fv, fcell);
  int iy_next = __ldg(year_index);
#ifdef H9_CYCLE_BUDGET
  Ticks tk;
#pragma unroll
  for (int k = 0; k < kTickSegs; ++k) tk.acc[k] = 0u;
  unsigned t_all0;
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

    /* the day's first fault is kept in three registers and recorded once per day: one warp per
     * scheduler has nothing to hide the wait for the step's last result behind */
    uint32_t ft_day = 0u, ft_first = 0u;
    int ft_ns = 0;
    float ft_imb = 0.0f;
    for (int ns = 0; ns < g.nisurf; ++ns) { /* HYBRID9.f90:193-211 */
      StepOut so;
      /* per warp and sub-step: the step without any water-table-in-column code when every cell
       * of the warp has a deep table (uniform branch), else the general straight-line step */
      const bool all_deep = __all_sync(kFullMask, !(s.zwt <= g.zim[NL]));
      const uint32_t ft = all_deep ? hydrology_step_pair<kStepAllDeep>(g, gd, pg, tbl, day, s, so, h, lane H9_TICKS_ARG)
                                   : hydrology_step_pair<kStepGeneral>(g, gd, pg, tbl, day, s, so, h, lane H9_TICKS_ARG);
      const bool first = (ft != 0u) && (ft_day == 0u);
      ft_first = first ? ft : ft_first;
      ft_ns = first ? ns : ft_ns;
      ft_imb = first ? so.imbalance : ft_imb;
      ft_day |= ft;
      /* H9_OPT_REAL_EVAP: evap_sum stays 0 otherwise, as in the reference (HYBRID9.f90:137,276) */
      evap_sum = fmaf(evap_on, so.qflx_evap_grnd + so.qflx_tran_veg_col, evap_sum);
    }
    if (ft_day && writer) {
      record_fault(a, c, sticky, ft_first, step0 + (unsigned long long)d * g.nisurf + ft_ns, ft_imb);
      if ((sticky | ft_day) != sticky) { /* further bits raised later in the day */
        atomicOr(a.any_fault, ft_day);
        sticky |= ft_day;
      }
    }

    /* GROW (HYBRID9.f90:217) on the gathered column: reads smp and rootr, rewrites rootr */
    float h2o_full[NL];
    pair_gather(s.smp, h, pmask, gs.smp);
    pair_gather(s.h2o, h, pmask, h2o_full);
    {
      const float4 r4 = tbl.rootr4();
      const float own[NH] = {r4.x, r4.y, r4.z, r4.w};
      pair_gather(own, h, pmask, gs.rootr);
    }
    GrowOut go;
    grow_day<MathFast>(g, tas, gs, go);
    tbl.set_rootr(gs.rootr, h);

    /* daily accumulators HYBRID9.f90:242-253 (even lane) */
    nt += 1;
    const bool close = (d + 1 == ndays) || (iy_next != iy);
    if (writer) {
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
        theta_sum[i] += fmaxf(h2o_full[i], 1.0E-6f) * g.rdzw[i + 1];
        h2osoi_sum_total += h2o_full[i];
      }
      a.npp_sum[c] = npp_sum;
      a.plant_mass_sum[c] = plant_mass_sum;
      a.h2osoi_sum_total[c] = h2osoi_sum_total;
      store8(a.theta_sum, c, theta_sum);

      /* year-end means HYBRID9.f90:263-291 (also at the end of a batch) */
      if (close && iy >= 1 && iy <= a.nyr) {
        float* out = a.annual + ((size_t)(iy - 1) * kAnnualPlanes) * a.ncs + c;
        const float rnt = MathFast::rcp((float)nt);
        out[0] = npp_sum;
        out[(size_t)1 * a.ncs] = plant_mass_sum * rnt;
        out[(size_t)2 * a.ncs] = s.rnf_sum * MathFast::rcp((float)(nt * g.nisurf));
        out[(size_t)3 * a.ncs] = a.real_evap ? evap_sum * MathFast::rcp((float)(nt * g.nisurf)) : 0.0f;
        out[(size_t)4 * a.ncs] = h2osoi_sum_total * rnt;
#pragma unroll
        for (int i = 0; i < NL; ++i) out[(size_t)(5 + i) * a.ncs] = theta_sum[i] * rnt;
      }
    }
  }

#ifdef H9_CYCLE_BUDGET
  if (lane == 0) {
    unsigned t_w;
    asm volatile("mov.u32 %0, %%clock;" : "=r"(t_w)::"memory");
    const int wg = (blockIdx.x * BLOCK + threadIdx.x) >> 5;
    if (wg < 4096) {
      g_pair_warp[0][wg] = tk.acc[0]; /* block A, all-deep */
      g_pair_warp[1][wg] = tk.acc[1]; /* tail, all-deep */
      g_pair_warp[2][wg] = tk.acc[5]; /* block A, general */
      g_pair_warp[3][wg] = tk.acc[2]; /* tail, general */
      g_pair_warp[4][wg] = tk.acc[3]; /* sub-steps all-deep */
      g_pair_warp[5][wg] = tk.acc[4]; /* sub-steps general */
      g_pair_warp[6][wg] = t_w - t_all0;
    }
  }
#endif
  if (valid) { /* each lane stores its half of the column, the even lane the scalars */
    store_half(a.h2o, c, h, s.h2o);
    store_half(a.smp, c, h, s.smp);
  }
  if (writer) {
    a.zwt[c] = s.zwt;
    a.wa[c] = s.wa;
    a.rnf_sum[c] = s.rnf_sum;
    store_grow_state(a, c, gs);
    a.fault[c] = sticky;
    if (a.real_evap) a.evap_sum[c] = evap_sum;
  }
}

__global__ void __launch_bounds__(128)
hydrology_step_kernel_pair(const DevArrays a, const __grid_constant__ Geo g, const ForcingView fv,
                           unsigned long long step0, const StepDiagArrays diag) {
  extern __shared__ float4 smem[];
  __shared__ float s_geo[kGeoDynFloats];
  GeoDyn::fill(s_geo, g);
  const GeoDyn gd{s_geo};
  const int lane = threadIdx.x & 31;
  const int h = threadIdx.x & 1;
  const int cell = (blockIdx.x * blockDim.x + threadIdx.x) >> 1;
  const bool valid = cell < a.nc;
  const int c = valid ? cell : a.nc - 1;
  const bool writer = valid && h == 0;
  const unsigned pmask = 3u << (lane & ~1);
  const PairTable<128> tbl{smem + threadIdx.x};
  PairGeo pg;
  pg.init(g, h);
  Params p;
  State gs;
  load_params(a, c, p);
  load_state(a, c, gs);
  tbl.init(g, p, gs.rootr, h);
  PairState s;
  load_half(a.h2o, c, h, s.h2o);
  load_half(a.smp, c, h, s.smp);
  s.zwt = gs.zwt;
  s.wa = gs.wa;
  s.rnf_sum = gs.rnf_sum;
  uint32_t sticky = a.fault[c];
  const Forcing f = load_forcing(fv, (size_t)c);
  DayFast day;
  float tas;
  day_setup_fast(g, f, gs.lai, gs.lai_litter, day, tas);
  StepOut so;
  const bool all_deep = __all_sync(kFullMask, !(s.zwt <= g.zim[NL]));
#ifdef H9_CYCLE_BUDGET
  Ticks tk;
#endif
  const uint32_t ft = all_deep ? hydrology_step_pair<kStepAllDeep>(g, gd, pg, tbl, day, s, so, h, lane H9_TICKS_ARG)
                               : hydrology_step_pair<kStepGeneral>(g, gd, pg, tbl, day, s, so, h, lane H9_TICKS_ARG);
  if (ft && writer) record_fault(a, c, sticky, ft, step0, so.imbalance);
  float h2o_full[NL];
  pair_gather(s.h2o, h, pmask, h2o_full);
  if (valid) {
    store_half(a.h2o, c, h, s.h2o);
    store_half(a.smp, c, h, s.smp);
  }
  if (!writer) return;
  a.zwt[c] = s.zwt;
  a.wa[c] = s.wa;
  a.rnf_sum[c] = s.rnf_sum;
  a.fault[c] = sticky;
  if (diag.theta) {
    float th[NL];
#pragma unroll
    for (int i = 0; i < NL; ++i) th[i] = fmaxf(h2o_full[i], 1.0E-6f) * g.rdzw[i + 1];
    store8(diag.theta, c, th);
  }
  if (diag.qflx_tran_veg_col) diag.qflx_tran_veg_col[c] = so.qflx_tran_veg_col;
  if (diag.qflx_evap_grnd) diag.qflx_evap_grnd[c] = so.qflx_evap_grnd;
  if (diag.rnf_inc) diag.rnf_inc[c] = so.rnf_inc;
  if (diag.w_imbalance) diag.w_imbalance[c] = so.imbalance;
  if (diag.jwt) diag.jwt[c] = so.jwt;
}

int launch_days_pair(void* stream, const DevArrays& a, const Geo& g, int ndays,
                     const int32_t* d_year_index, const ForcingView& fv, int cur_year, int nt,
                     unsigned long long step0, const int32_t* d_cell_index) {
  cudaStream_t st = (cudaStream_t)stream;
  if (a.nc <= 0 || ndays <= 0) return 0;
  constexpr int BS = 128; /* 4 warps = one per scheduler, 64 cells per block */
  const size_t shm = (size_t)kPairFloatsPerLane * BS * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(days_kernel_pair<BS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
  if (e != cudaSuccess) return (int)e;
  const int cpb = BS / 2;
  days_kernel_pair<BS><<<(a.nc + cpb - 1) / cpb, BS, shm, st>>>(a, g, ndays, d_year_index, fv, cur_year,
                                                               nt, step0, d_cell_index);
  return (int)cudaGetLastError();
}

int launch_hydrology_step_pair(void* stream, const DevArrays& a, const Geo& g, const ForcingView& fv,
                               unsigned long long step0, const StepDiagArrays& diag) {
  if (a.nc <= 0) return 0;
  const size_t shm = (size_t)kPairFloatsPerLane * 128 * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(hydrology_step_kernel_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
  if (e != cudaSuccess) return (int)e;
  hydrology_step_kernel_pair<<<(a.nc + 63) / 64, 128, shm, (cudaStream_t)stream>>>(a, g, fv, step0, diag);
  return (int)cudaGetLastError();
}

} /* namespace h9 */
