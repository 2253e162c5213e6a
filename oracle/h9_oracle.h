/*
 * h9_oracle.h -- CPU restatement of HYBRID9's HYDROLOGY + GROW + driver loop.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (hybrid9_b200/,
 * libh9gpu.so) may link, import or execute anything under oracle/.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs use it, and there only as the checker / CPU baseline.
 *
 * PINNING.  The reference ships no tests, golden vectors or fixtures
 * and the image has no Fortran compiler, MPI or netCDF, so the reference program itself
 * cannot be built.  What pins this restatement instead is the reference's OWN source text,
 * compiled: oracle/f2cpp.py translates HYDROLOGY.f90, GROW.f90 and the relevant line ranges
 * of INIT.f90 / HYBRID9.f90 statement for statement into C++ (oracle/_ref/, built from
 * /root/reference where it lies, never copied) and tests/test_ref_vs_oracle.py holds this
 * file to it BIT FOR BIT: geometry, calendar, initial state, land mask, single HYDROLOGY calls
 * with the water table in every layer, GROW, multi-day runs with and without the smp leak, a
 * whole 3652-day decade through the verbatim loop nest, and the STOP record.  Outputs of that
 * build are committed as tests/golden/h9_ref_golden_v1.npz.  Still not covered: the libm of the
 * compiler the reference was actually built with (ifort's SVML); DESIGN.md section 4.
 *
 * Build: see oracle/Makefile.  `real` is float by default (the reference is
 * default-kind REAL throughout) and double with -DH9O_DOUBLE (noise floor).
 */
#ifndef H9_ORACLE_H
#define H9_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef H9O_DOUBLE
typedef double h9o_real;
#else
typedef float h9o_real;
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct h9o_ctx h9o_ctx;

/* per-step diagnostics of one HYDROLOGY call (locals of HYDROLOGY.f90) */
typedef struct h9o_step_diag {
  h9o_real theta[8];
  h9o_real qflx_tran_veg_col, qflx_evap_grnd, qflx_surf, rsub_top, qflx_rsub_sat;
  h9o_real qflx_infl, qcharge, fsat, beta, rsc, w0, w1;
  h9o_real rnf_inc; /* (qflx_surf + rsub_top)*dt as added to rnf_sum, :1282-1283 */
  int32_t jwt_soilwater; /* jwt of :499-508 */
  int32_t jwt_final;     /* jwt after Drainage, the one :1166 tests */
  uint32_t fault;        /* H9_FAULT_* bits */
} h9o_step_diag;

/* The API mirrors include/h9gpu.h one to one (same argument meaning and array
 * layouts) so that a parity test issues the same calls to both sides. */
int h9o_create(h9o_ctx** ctx);
int h9o_destroy(h9o_ctx* ctx);
int h9o_configure(h9o_ctx* ctx, int lon_c, int lat_c, int nisurf, const h9o_real zi[10], int nyr);
int h9o_set_soil(h9o_ctx* ctx, const int32_t* soil_tex, const h9o_real* theta_s,
                 const h9o_real* hksat, const h9o_real* bsw, const h9o_real* psi_s,
                 const h9o_real* fmax);
int64_t h9o_num_land(const h9o_ctx* ctx);
int h9o_get_land_index(const h9o_ctx* ctx, int32_t* cell_xy);

/* INIT.f90:707-811 for every land cell of the block (writes the ctx state). */
int h9o_init_state(h9o_ctx* ctx);

int h9o_set_state(h9o_ctx* ctx, const h9o_real* h2osoi_liq, const h9o_real* zwt,
                  const h9o_real* wa, const h9o_real* lai, const h9o_real* lai_litter,
                  const h9o_real* plant_mass, const h9o_real* plant_foliage_mass,
                  const h9o_real* plant_length, const h9o_real* rdepth, const h9o_real* rootr_col,
                  const int32_t* nplants, const h9o_real* smp);
int h9o_get_state(h9o_ctx* ctx, h9o_real* h2osoi_liq, h9o_real* zwt, h9o_real* wa, h9o_real* lai,
                  h9o_real* lai_litter, h9o_real* plant_mass, h9o_real* plant_foliage_mass,
                  h9o_real* plant_length, h9o_real* rdepth, h9o_real* rootr_col, int32_t* nplants,
                  h9o_real* smp);

/* loop_order: 0 = cell-outer / time-inner (the reference's HYBRID9.f90:120-295),
 *             1 = time-outer / cell-inner (what a GPU does).
 * smp_leak:   1 = keep `smp` as ONE scratch vector shared by all cells like
 *             SHARED.f90:198 (only meaningful with loop_order 0), 0 = per-cell.
 * nthreads:   >1 splits the land cells over that many threads (cell-outer only;
 *             used by the CPU baseline; cells are independent). */
int h9o_set_options(h9o_ctx* ctx, int loop_order, int smp_leak, int nthreads);

/* 0 (default): axy_evap == 0 like the reference (evap_sum is never accumulated,
 * HYBRID9.f90:137,276); 1: annual mean of qflx_evap_grnd + qflx_tran_veg_col (mm/s) */
int h9o_set_real_evap(h9o_ctx* ctx, int on);

int h9o_run_days(h9o_ctx* ctx, int ndays, const int32_t* year_index_of_day, const h9o_real* tas,
                 const h9o_real* rlds, const h9o_real* rsds, const h9o_real* huss,
                 const h9o_real* ps, const h9o_real* pr, const h9o_real* rhs);

int h9o_get_annual(h9o_ctx* ctx, int iyr, h9o_real* axy_npp, h9o_real* axy_plant_mass,
                   h9o_real* axy_rnf, h9o_real* axy_evap, h9o_real* axy_theta_total,
                   h9o_real* axy_theta);

/* first fault in reference iteration order of the last run (x,y 1-based) */
int h9o_get_fault(h9o_ctx* ctx, uint32_t* any, uint32_t* code, int32_t* x, int32_t* y,
                  int32_t* day, int32_t* substep, h9o_real* imbalance, int64_t* n_faulted);
int h9o_clear_fault(h9o_ctx* ctx);

int h9o_hydrology_step(h9o_ctx* ctx, const h9o_real* tas, const h9o_real* rlds,
                       const h9o_real* rsds, const h9o_real* huss, const h9o_real* ps,
                       const h9o_real* pr, const h9o_real* rhs, h9o_real* theta,
                       h9o_real* qflx_tran_veg_col, h9o_real* qflx_evap_grnd, h9o_real* rnf_inc,
                       h9o_real* w_imbalance, int32_t* jwt);
int h9o_grow_day(h9o_ctx* ctx, const h9o_real* tas, h9o_real* npp, h9o_real* w_i, h9o_real* fT);

/* full per-cell diagnostics of the last h9o_hydrology_step for block cell (x,y), 1-based */
int h9o_last_step_diag(h9o_ctx* ctx, int x, int y, h9o_step_diag* out);

/* INIT.f90:573-633 for one layer (see h9_regrid_soil_layer in include/h9gpu.h) */
int h9o_regrid_soil_layer(int lon_c, int lat_c, int layer, const h9o_real* theta_s_in,
                          const h9o_real* k_s_in, const h9o_real* lambda_in, const h9o_real* psi_s_in,
                          h9o_real* theta_s, h9o_real* hksat, h9o_real* bsw, h9o_real* psi_s);

/* geometry as INIT.f90:214,252-257 computes it: dz[1..9], zc[1..9] (index 0 unused), dt */
int h9o_get_geometry(const h9o_ctx* ctx, h9o_real dz[10], h9o_real zc[10], h9o_real* dt);

/* time_BOY (INIT.f90:844-859): day number (1 = 1 Jan 1860) of 1 Jan of `year`, 1860..2300 */
int h9o_time_boy(int year);

int h9o_sizeof_real(void);

#ifdef __cplusplus
}
#endif
#endif
