/*
 * h9_kernels.cuh -- the time-stepping kernels, templated on the arithmetic
 * policy.  Included by h9_kernels_exact.cu (built -fmad=false) and
 * h9_kernels_fast.cu.
 *
 *   days_kernel            K3: for each land cell (one thread), for each day of
 *                          the batch: forcing derivation (HYBRID9.f90:168-184),
 *                          NISURF x HYDROLOGY (:193-211), GROW (:217), the
 *                          daily/annual accumulators (:235-254) and the year-end
 *                          means (:263-291).  State lives in registers for the
 *                          whole batch; HBM traffic is the forcing stream plus
 *                          one state load/store per launch.
 *   hydrology_step_kernel  K1: one HYDROLOGY call for all cells (parity target)
 *   grow_day_kernel        K2: one GROW call for all cells (parity target)
 */
#ifndef H9_KERNELS_CUH
#define H9_KERNELS_CUH

#include <cuda_runtime.h>

#include "h9_device.h"

namespace h9 {

__device__ __forceinline__ void load8(const float* base, int c, float (&v)[NL]) {
  const float4* q = reinterpret_cast<const float4*>(base) + 2 * (size_t)c;
  const float4 a = q[0], b = q[1];
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

__device__ __forceinline__ void store8(float* base, int c, const float (&v)[NL]) {
  float4* q = reinterpret_cast<float4*>(base) + 2 * (size_t)c;
  q[0] = make_float4(v[0], v[1], v[2], v[3]);
  q[1] = make_float4(v[4], v[5], v[6], v[7]);
}

__device__ __forceinline__ void load_params(const DevArrays& a, int c, Params& p) {
  H9_ASSERT(c >= 0 && c < a.ncs);
  load8(a.theta_s, c, p.theta_s);
  load8(a.hksat, c, p.hksat);
  load8(a.bsw, c, p.bsw);
  load8(a.psi_s, c, p.psi_s);
  p.fmax = a.fmax[c];
}

__device__ __forceinline__ void load_state(const DevArrays& a, int c, State& s) {
  load8(a.h2o, c, s.h2o);
  load8(a.smp, c, s.smp);
  load8(a.rootr, c, s.rootr);
  s.zwt = a.zwt[c];
  s.wa = a.wa[c];
  s.lai = a.lai[c];
  s.lai_litter = a.lai_litter[c];
  s.plant_mass = a.plant_mass[c];
  s.plant_foliage_mass = a.plant_foliage_mass[c];
  s.plant_length = a.plant_length[c];
  s.rdepth = a.rdepth[c];
  s.rnf_sum = a.rnf_sum[c];
  s.nplants = a.nplants[c];
}

__device__ __forceinline__ void store_hydro_state(const DevArrays& a, int c, const State& s) {
  H9_ASSERT(c >= 0 && c < a.nc);
  store8(a.h2o, c, s.h2o);
  store8(a.smp, c, s.smp);
  a.zwt[c] = s.zwt;
  a.wa[c] = s.wa;
  a.rnf_sum[c] = s.rnf_sum;
}

__device__ __forceinline__ void store_grow_state(const DevArrays& a, int c, const State& s) {
  store8(a.rootr, c, s.rootr);
  a.lai[c] = s.lai;
  a.lai_litter[c] = s.lai_litter;
  a.plant_mass[c] = s.plant_mass;
  a.plant_foliage_mass[c] = s.plant_foliage_mass;
  a.plant_length[c] = s.plant_length;
  a.rdepth[c] = s.rdepth;
}

__device__ __forceinline__ Forcing load_forcing(const ForcingView& fv, size_t off) {
  H9_ASSERT(off < fv.limit);
  Forcing r;
  r.tas = __ldg(fv.plane[0] + off);
  r.rlds = __ldg(fv.plane[1] + off);
  r.rsds = __ldg(fv.plane[2] + off);
  r.huss = __ldg(fv.plane[3] + off);
  r.ps = __ldg(fv.plane[4] + off);
  r.pr = __ldg(fv.plane[5] + off);
  r.rhs = __ldg(fv.plane[6] + off);
  return r;
}

/* first fault of a cell: the record the reference would print before STOP */
__device__ __forceinline__ void record_fault(const DevArrays& a, int c, uint32_t& sticky,
                                             uint32_t code, unsigned long long step, float imb) {
  if (sticky == 0u) {
    a.first_code[c] = code;
    a.first_step[c] = step;
    a.first_imb[c] = imb;
    atomicMin(a.first_key, (step << 32) | (unsigned long long)(unsigned)c);
  }
  if ((sticky | code) != sticky) atomicOr(a.any_fault, code);
  sticky |= code;
}

template <class M, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB)
days_kernel(const DevArrays a, const __grid_constant__ Geo g, int ndays,
            const int32_t* __restrict__ year_index, const ForcingView fv, int cur_year, int nt,
            unsigned long long step0, const int32_t* __restrict__ cell_index) {
  const int c = blockIdx.x * BLOCK + threadIdx.x;
  if (c >= a.nc) return;

  Params p;
  State s;
  load_params(a, c, p);
  load_state(a, c, s);
  float npp_sum = a.npp_sum[c];
  float plant_mass_sum = a.plant_mass_sum[c];
  float h2osoi_sum_total = a.h2osoi_sum_total[c];
  float theta_sum[NL];
  load8(a.theta_sum, c, theta_sum);
  uint32_t sticky = a.fault[c];
  float evap_sum = a.real_evap ? a.evap_sum[c] : 0.0f;

  /* compact forcing [day][7][ncs] (cell_index == nullptr) or the grid tile as the host holds
   * it, gathered through the land index */
  const size_t fcell = (size_t)(cell_index ? __ldg(cell_index + c) : c);
  Forcing fnext = load_forcing(fv, fcell);
  int iy_next = __ldg(year_index);

  for (int d = 0; d < ndays; ++d) {
    const Forcing f = fnext;
    const int iy = iy_next;
    if (d + 1 < ndays) { /* prefetch the next day's forcing behind this day's 48 sub-steps */
      fnext = load_forcing(fv, fcell + (size_t)(d + 1) * fv.day_stride);
      iy_next = __ldg(year_index + d + 1);
    }
    if (iy != cur_year) { /* HYBRID9.f90:134-146 */
      cur_year = iy;
      nt = 0;
      npp_sum = 0.0f;
      plant_mass_sum = 0.0f;
      s.rnf_sum = 0.0f;
      evap_sum = 0.0f;
      h2osoi_sum_total = 0.0f;
#pragma unroll
      for (int i = 0; i < NL; ++i) theta_sum[i] = 0.0f;
    }

    Day day;
    day_setup<M>(g, f, s.lai, s.lai_litter, day);

    for (int ns = 0; ns < g.nisurf; ++ns) { /* HYBRID9.f90:193-211 */
      StepOut so;
      const uint32_t ft = hydrology_step<M>(g, p, day, s, so);
      if (ft)
        record_fault(a, c, sticky, ft, step0 + (unsigned long long)d * g.nisurf + ns,
                     so.imbalance);
      if (a.real_evap) evap_sum = evap_sum + (so.qflx_evap_grnd + so.qflx_tran_veg_col);
    }

    GrowOut go;
    grow_day<M>(g, day.tas, s, go); /* HYBRID9.f90:217 */

    /* HYBRID9.f90:242-253 */
    if (s.nplants >= 1) plant_mass_sum = plant_mass_sum + s.plant_mass;
    npp_sum = npp_sum + go.npp;
#pragma unroll
    for (int i = 0; i < NL; ++i) {
      theta_sum[i] = theta_sum[i] + theta_diag<M>(g, s.h2o[i], i);
      h2osoi_sum_total = h2osoi_sum_total + s.h2o[i];
    }
    nt += 1;

    /* HYBRID9.f90:263-291 at the last day of a year (or of the batch, so that a
     * year split over several calls is always up to date) */
    const bool close = (d + 1 == ndays) || (iy_next != iy);
    if (close && iy >= 1 && iy <= a.nyr) {
      float* out = a.annual + ((size_t)(iy - 1) * kAnnualPlanes) * a.ncs + c;
      const float fnt = (float)nt;
      out[0] = npp_sum;
      out[(size_t)1 * a.ncs] = M::div(plant_mass_sum, fnt);
      out[(size_t)2 * a.ncs] = M::div(s.rnf_sum, (float)(nt * g.nisurf));
      /* the reference's evap_sum never accumulates (axy_evap == 0) unless H9_OPT_REAL_EVAP */
      out[(size_t)3 * a.ncs] = M::div(a.real_evap ? evap_sum : 0.0f, (float)(nt * g.nisurf));
      out[(size_t)4 * a.ncs] = M::div(h2osoi_sum_total, fnt);
#pragma unroll
      for (int i = 0; i < NL; ++i) out[(size_t)(5 + i) * a.ncs] = M::div(theta_sum[i], fnt);
    }
  }

  store_hydro_state(a, c, s);
  store_grow_state(a, c, s);
  a.npp_sum[c] = npp_sum;
  a.plant_mass_sum[c] = plant_mass_sum;
  a.h2osoi_sum_total[c] = h2osoi_sum_total;
  store8(a.theta_sum, c, theta_sum);
  a.fault[c] = sticky;
  if (a.real_evap) a.evap_sum[c] = evap_sum;
}

template <class M>
__global__ void __launch_bounds__(128)
hydrology_step_kernel(const DevArrays a, const __grid_constant__ Geo g, const ForcingView fv,
                      unsigned long long step0, const StepDiagArrays diag) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.nc) return;
  Params p;
  State s;
  load_params(a, c, p);
  load_state(a, c, s);
  uint32_t sticky = a.fault[c];
  const Forcing f = load_forcing(fv, (size_t)c);
  Day day;
  day_setup<M>(g, f, s.lai, s.lai_litter, day);
  StepOut so;
  const uint32_t ft = hydrology_step<M>(g, p, day, s, so);
  if (ft) record_fault(a, c, sticky, ft, step0, so.imbalance);
  store_hydro_state(a, c, s);
  a.fault[c] = sticky;
  if (diag.theta) {
    float th[NL];
#pragma unroll
    for (int i = 0; i < NL; ++i) th[i] = theta_diag<M>(g, s.h2o[i], i);
    store8(diag.theta, c, th);
  }
  if (diag.qflx_tran_veg_col) diag.qflx_tran_veg_col[c] = so.qflx_tran_veg_col;
  if (diag.qflx_evap_grnd) diag.qflx_evap_grnd[c] = so.qflx_evap_grnd;
  if (diag.rnf_inc) diag.rnf_inc[c] = so.rnf_inc;
  if (diag.w_imbalance) diag.w_imbalance[c] = so.imbalance;
  if (diag.jwt) diag.jwt[c] = so.jwt;
}

template <class M>
__global__ void __launch_bounds__(128)
grow_day_kernel(const DevArrays a, const __grid_constant__ Geo g, const float* __restrict__ tas,
                const GrowDiagArrays diag) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.nc) return;
  State s;
  load_state(a, c, s);
  GrowOut go;
  grow_day<M>(g, __ldg(tas + c), s, go);
  store_grow_state(a, c, s);
  if (diag.npp) diag.npp[c] = go.npp;
  if (diag.w_i) diag.w_i[c] = go.w_i;
  if (diag.fT) diag.fT[c] = go.fT;
}

/* Which build of the generic (exact-mode) day kernel: the one capped at 128 registers runs a grid
 * of more than 8 warps per SM in one wave and is 24 % faster there (616 vs 810 ms per simulated
 * 0.5 deg year) although it spills 424 bytes per thread; a small shard keeps the uncapped one.
 * block 1064 = automatic, other values as documented in include/h9gpu.h. */
inline bool days_exact_capped(int nc, int block) {
  if (block < 1000 || block >= 2000) return false;
  if (block != 1064) return true;
  int dev = 0, nsm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  return nc > nsm * 8 * 32;
}
template <class M>
int launch_days_t(void* stream, const DevArrays& a, const Geo& g, int ndays,
                  const int32_t* d_year_index, const ForcingView& fv, int cur_year, int nt,
                  unsigned long long step0, int block, const int32_t* d_cell_index) {
  cudaStream_t st = (cudaStream_t)stream;
  if (a.nc <= 0 || ndays <= 0) return 0;
  /* block: threads per block (32/64/128) + 1000 when the kernel variant capped at 128
   * registers per thread (16 warps per SM resident: the whole 0.5 deg grid in one wave) is
   * wanted instead of the uncapped one (255 registers, 8 warps per SM) */
  const bool capped = days_exact_capped(a.nc, block);
  const int bs = block >= 2000 ? 64 : block % 1000;
#define H9_LAUNCH(BS, MINB)                                                                      \
  days_kernel<M, BS, MINB><<<(a.nc + BS - 1) / BS, BS, 0, st>>>(                                 \
      a, g, ndays, d_year_index, fv, cur_year, nt, step0, d_cell_index)
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

template <class M>
int launch_hydrology_step_t(void* stream, const DevArrays& a, const Geo& g,
                            const ForcingView& fv, unsigned long long step0,
                            const StepDiagArrays& diag) {
  if (a.nc <= 0) return 0;
  hydrology_step_kernel<M><<<(a.nc + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      a, g, fv, step0, diag);
  return (int)cudaGetLastError();
}

template <class M>
int launch_grow_day_t(void* stream, const DevArrays& a, const Geo& g, const float* d_tas,
                      const GrowDiagArrays& diag) {
  if (a.nc <= 0) return 0;
  grow_day_kernel<M><<<(a.nc + 127) / 128, 128, 0, (cudaStream_t)stream>>>(a, g, d_tas, diag);
  return (int)cudaGetLastError();
}

#define H9_DEFINE_LAUNCHERS(SUFFIX, POLICY)                                                      \
  int launch_days_##SUFFIX(void* stream, const DevArrays& a, const Geo& g, int ndays,            \
                           const int32_t* d_year_index, const ForcingView& fv,                   \
                           int cur_year, int nt,                                                 \
                           unsigned long long step0, int block,                                  \
                           const int32_t* d_cell_index) {                                        \
    return launch_days_t<POLICY>(stream, a, g, ndays, d_year_index, fv, cur_year, nt, step0,     \
                                 block, d_cell_index);                                           \
  }                                                                                              \
  int launch_hydrology_step_##SUFFIX(void* stream, const DevArrays& a, const Geo& g,             \
                                     const ForcingView& fv,                                      \
                                     unsigned long long step0, const StepDiagArrays& diag) {     \
    return launch_hydrology_step_t<POLICY>(stream, a, g, fv, step0, diag);                       \
  }                                                                                              \
  int launch_grow_day_##SUFFIX(void* stream, const DevArrays& a, const Geo& g,                   \
                               const float* d_tas, const GrowDiagArrays& diag) {                 \
    return launch_grow_day_t<POLICY>(stream, a, g, d_tas, diag);                                 \
  }

} /* namespace h9 */
#endif
