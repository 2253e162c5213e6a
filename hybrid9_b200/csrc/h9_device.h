/*
 * h9_device.h -- device data layout shared by the kernels and the C-ABI layer.
 *
 * HBM layout (DESIGN.md "Data layout"): land cells are compacted in the
 * reference's iteration order (HYBRID9.f90:120-123) and padded to a multiple of
 * 128 (`ncs`, the cell stride).  Per-layer fields are cell-major with the 8 soil
 * layers innermost ([cell][8] == two float4 per cell, the Fortran (8,x,y) order
 * of SHARED.f90:398-459 after compaction); per-cell scalars are flat [ncs]
 * planes; forcing is day-major [day][7][ncs]; annual means are [year][13][ncs].
 */
#ifndef H9_DEVICE_H
#define H9_DEVICE_H

#include <stddef.h>
#include <stdint.h>

#include "h9_physics.h"

namespace h9 {

constexpr int kCellPad = 128;      /* ncs is a multiple of this */
constexpr int kAnnualPlanes = 13;  /* npp, plant_mass, rnf, evap, theta_total, theta(1..8) */
constexpr int kForcingPlanes = 7;  /* tas, rlds, rsds, huss, ps, pr, rhs (READ_PGF.f90 order) */

/* the seven forcing fields of one batch of days: plane[k][d*day_stride + cell].  The planes
 * may be slices of one compact device buffer, of a staged grid tile, or the host's own
 * page-locked arrays mapped into the device address space (read in place over PCIe). */
struct ForcingView {
  const float* plane[kForcingPlanes];
  size_t day_stride;
  size_t limit; /* elements of one plane's view; read by -DH9_BOUNDS_CHECK builds only */
};

struct DevArrays {
  int nc;  /* land cells of this ctx */
  int ncs; /* padded cell stride */
  int nyr;
  /* state, [ncs][8] */
  float *h2o, *smp, *rootr;
  /* parameters, [ncs][8] */
  const float *theta_s, *hksat, *bsw, *psi_s;
  /* per-cell scalars, [ncs] */
  float *zwt, *wa, *lai, *lai_litter, *plant_mass, *plant_foliage_mass, *plant_length, *rdepth;
  const float* fmax;
  int32_t* nplants;
  /* annual accumulators of HYBRID9.f90:134-146 (per cell) */
  float *rnf_sum, *npp_sum, *plant_mass_sum, *h2osoi_sum_total; /* [ncs] */
  float* evap_sum;   /* [ncs], only accumulated when real_evap != 0 (H9_OPT_REAL_EVAP) */
  int real_evap;
  float* theta_sum;                                              /* [ncs][8] */
  /* annual means, HYBRID9.f90:263-291: [nyr][13][ncs] */
  float* annual;
  /* faults */
  uint32_t* fault;               /* sticky word per cell */
  uint32_t* first_code;          /* fault bits of the cell's first faulting step */
  unsigned long long* first_step; /* global sub-step index (0-based) of that step */
  float* first_imb;              /* w1-w0 of that step */
  unsigned long long* first_key; /* min over cells of (step << 32 | cell); ~0 = none */
  uint32_t* any_fault;           /* OR over all cells */
  int budget_cell;               /* -DH9_CYCLE_BUDGET builds only: the cell whose cycles are recorded */
};

/* optional per-cell outputs of the fine-grained sub-step entry, compact [ncs] */
struct StepDiagArrays {
  float* theta; /* [ncs][8] */
  float *qflx_tran_veg_col, *qflx_evap_grnd, *rnf_inc, *w_imbalance;
  int32_t* jwt;
};

struct GrowDiagArrays {
  float *npp, *w_i, *fT;
};

/* launchers; one set per arithmetic mode, defined in h9_kernels_{exact,fast}.cu.
 * All return the cudaError_t of the launch as int. */
#define H9_DECLARE_LAUNCHERS(SUFFIX)                                                             \
  int launch_days_##SUFFIX(void* stream, const DevArrays& a, const Geo& g, int ndays,            \
                           const int32_t* d_year_index, const ForcingView& fv,                   \
                           int cur_year, int nt,                                                 \
                           unsigned long long step0, int block,                                  \
                           const int32_t* d_cell_index /* nullptr: compact forcing */);          \
  int launch_hydrology_step_##SUFFIX(void* stream, const DevArrays& a, const Geo& g,             \
                                     const ForcingView& fv,                                      \
                                     unsigned long long step0, const StepDiagArrays& diag);      \
  int launch_grow_day_##SUFFIX(void* stream, const DevArrays& a, const Geo& g,                   \
                               const float* d_tas, const GrowDiagArrays& diag);

H9_DECLARE_LAUNCHERS(exact)
H9_DECLARE_LAUNCHERS(fast)

/* fast mode, two lanes per land cell (h9_kernels_pair.cu): small shards */
int launch_days_pair(void* stream, const DevArrays& a, const Geo& g, int ndays,
                     const int32_t* d_year_index, const ForcingView& fv, int cur_year, int nt,
                     unsigned long long step0, const int32_t* d_cell_index);
int launch_hydrology_step_pair(void* stream, const DevArrays& a, const Geo& g, const ForcingView& fv,
                               unsigned long long step0, const StepDiagArrays& diag);
int launch_hydrology_step_fast_variant(void* stream, const DevArrays& a, const Geo& g,
                                       const ForcingView& fv, unsigned long long step0,
                                       const StepDiagArrays& diag, int block);
/* name of the day-kernel instantiation launch_days_fast picks for (nc, block) */
const char* days_variant_fast(int nc, int block);
const char* days_variant_exact(int nc, int block);

/* mode-independent kernels, h9_pack.cu */
int launch_pack_forcing(void* stream, const float* const planes[kForcingPlanes] /* each [ndays][ngrid] */,
                        int ndays, size_t ngrid, const int32_t* d_cell_xy, int nc, int ncs,
                        float* d_out /* [ndays][7][ncs] */);
int launch_clear_u32(void* stream, uint32_t* p, size_t n, uint32_t v);
/* N4: 60x60 block means of the four 30-arc-second soil fields for `rows` coarse rows */
int launch_regrid_soil(void* stream, const float* d_in /* [4][rows*60][lon_c*60] */, int lon_c, int rows,
                       float* d_out /* [4][rows][lon_c]: theta_s, hksat, bsw, psi_s */);

} /* namespace h9 */
#endif
