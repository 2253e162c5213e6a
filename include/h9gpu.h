/*
 * h9gpu.h -- C ABI of libh9gpu, the B200-native replacement for HYBRID9's
 * per-grid-cell time-stepping hot path.
 *
 * What this boundary replaces (all citations relative to /root/reference/SOURCE):
 *
 *   The reference has no operator/plugin API: HYDROLOGY (HYDROLOGY.f90:2) and
 *   GROW (GROW.f90:2) are argument-less subroutines that work on the module
 *   globals of SHARED/CONTROL for ONE cell and ONE step, called from the
 *   6-deep loop nest of PROGRAM H9 (HYBRID9.f90:120-295, calls at :203 and
 *   :217).  A GPU needs all cells per step, so the drop-in boundary is that
 *   loop nest: the Fortran host keeps CONTROL/INIT/SHARED, driver.txt,
 *   READ_PGF and the WRITE_NET_CDF_* writers, and replaces the body of its
 *   decade loop by the calls below (see INTEGRATION.md for the ISO_C_BINDING
 *   interface module).
 *
 * Array conventions: every host pointer is the dense data of the Fortran
 * array named in the comment, column-major, exactly as the reference holds
 * it (C_LOC of the allocatable).  (8,lon_c,lat_c) means layer fastest.
 * lat index 1 is the northernmost row of the block (INIT.f90:145).
 * The caller owns all host pointers; the library copies what it needs
 * before returning.  One h9_ctx drives one GPU and is used by one host
 * thread (one MPI rank in the reference's 1-rank-per-block layout).
 *
 * Return value: 0 on success; >0 = a physics fault of the kind that makes the
 * reference STOP (see h9_fault); <0 = usage/CUDA error, text via
 * h9_last_error().
 */
#ifndef H9GPU_H
#define H9GPU_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct h9_ctx h9_ctx;

#define H9_NLAYERS 8 /* nsoil_layers_max, SHARED.f90:294 */
#define H9_NLEVGRND 9 /* Nlevgrnd, SHARED.f90:300 */

/* error codes (<0) */
#define H9_OK 0
#define H9_ERR_ARG -1
#define H9_ERR_CUDA -2
#define H9_ERR_STATE -3
#define H9_ERR_NOMEM -4
#define H9_ERR_NCCL -5

/* fault bits: the four conditions on which the reference STOPs */
#define H9_FAULT_TRIDIAG_PIVOT1 1u  /* bmx(1)==0      HYDROLOGY.f90:806-812  */
#define H9_FAULT_TRIDIAG_PIVOT2 2u  /* BET==0         HYDROLOGY.f90:818-825  */
#define H9_FAULT_RSUB_POSITIVE 4u   /* rsub_top_tot>0 HYDROLOGY.f90:1068-1071 */
#define H9_FAULT_WATER_IMBALANCE 8u /* |w1-w0|>0.1    HYDROLOGY.f90:1244-1274 */

/* arithmetic mode of the device kernels */
#define H9_MATH_EXACT 0 /* every + - * / is the IEEE operation the Fortran order implies (no FMA
                         * contraction, IEEE division); pow/exp/log by portable table-driven
                         * double-precision kernels (IEEE double ops + fma) whose float result is correctly
                         * rounded: the reference's arithmetic bit for bit, the verification mode */
#define H9_MATH_FAST 1  /* default: pow(a,b) = ex2.approx(b * lg2.approx(a)), exp by ex2.approx,
                         * 1/x by rcp.approx (MUFU), explicit FMAs, reciprocals shared and hoisted;
                         * deviates from EXACT at the FP32 rounding-noise level of the model */

typedef struct h9_fault {
  uint32_t any;       /* OR of the fault bits over all cells since the last clear */
  uint32_t code;      /* fault bits of the FIRST fault (earliest step, lowest cell) */
  int32_t x, y;       /* 1-based block indices of that cell, as in the reference's messages */
  int32_t day;        /* 1-based day since the ctx was created/reset (DiTIME analogue) */
  int32_t substep;    /* 1-based NS within the day */
  float imbalance;    /* w1-w0 (mm) of that step, for the "Water imbalance" message */
  int64_t n_faulted;  /* number of cells with a non-zero sticky fault word */
} h9_fault;

/* ---- life cycle ------------------------------------------------------- */

/* device_id < 0: use the current CUDA device. */
int h9_create(h9_ctx** ctx, int device_id);
int h9_destroy(h9_ctx* ctx);
const char* h9_last_error(const h9_ctx* ctx);

/* lon_c, lat_c: block shape (CONTROL.f90:49-50, INIT.f90:274-283);
 * nisurf: sub-steps per day (driver.txt:2, dt = 86400/NISURF INIT.f90:214);
 * zi[0..9]: layer interfaces in mm (driver.txt:17-26, INIT.f90:202-204);
 * nyr: number of years the annual diagnostics must hold (NYR, INIT.f90:286-290). */
int h9_configure(h9_ctx* ctx, int lon_c, int lat_c, int nisurf, const float zi[10], int nyr);

/* H9_MATH_EXACT or H9_MATH_FAST; may be switched between runs. */
int h9_set_math(h9_ctx* ctx, int mode);

/* Options.  H9_OPT_REAL_EVAP (default 0): the reference never accumulates evap_sum
 * (HYBRID9.f90:137,276), so axy_evap is identically 0 on land; 0 reproduces that, 1 makes
 * h9_get_annual return the annual mean of qflx_evap_grnd + qflx_tran_veg_col (mm/s) instead. */
#define H9_OPT_REAL_EVAP 1
int h9_set_option(h9_ctx* ctx, int option, int value);

/* Tuning knobs: days of forcing per pipeline tile of h9_run_days (default 8, env
 * H9_TILE_DAYS; no effect on results) and the launch shape of the fast-mode stepping
 * kernel, `block` (env H9_BLOCK):
 *   32 / 64 / 128     thread-per-cell kernel, that many threads per block;
 *   +1000             the same compiled for <=128 registers per thread (16 resident warps per
 *                     SM, so the 0.5 deg grid runs in one wave);
 *   4000              the two-lanes-per-cell kernel for small shards (layers 1..4 and 8..5 +
 *                     aquifer on a lane pair, tridiagonal solved from both ends);
 *   1064 (default)    automatic: two lanes per cell up to 16 cells x 4 schedulers x SMs
 *                     (9,472 on a B200; env H9_PAIR_MAX_CELLS overrides), thread per cell with
 *                     all registers up to 8 warps per SM (one 128-thread block per SM while
 *                     the shard has at most 128 cells per SM), the 128-register build above.
 * All thread-per-cell shapes give the same bits; the two-lanes-per-cell kernel orders the
 * tridiagonal solve and the column sums differently and agrees with them at rounding level
 * (FP32 noise floor of the model, tests/test_gpu_pair.py).  A run that must be bit-identical
 * whatever the shard size (shards == whole grid) pins the shape, e.g. block = 64.  <=0 keeps.
 * The exact-mode kernel is always thread per cell: threads per block as above, the 128-register
 * build with +1000 or, under 1064, for more than 8 warps per SM of cells; every shape gives the
 * same bits. */
int h9_set_tuning(h9_ctx* ctx, int tile_days, int block);
/* name of the stepping-kernel instantiation the next h9_run_days will launch */
const char* h9_kernel_variant(h9_ctx* ctx);

/* soil_tex (lon_c,lat_c) int32; theta_s,hksat,bsw,psi_s (8,lon_c,lat_c); fmax (lon_c,lat_c).
 * Builds the land mask with the reference's predicate and iteration order
 * (HYBRID9.f90:120-123) and uploads the compacted parameters. */
int h9_set_soil(h9_ctx* ctx, const int32_t* soil_tex, const float* theta_s, const float* hksat,
                const float* bsw, const float* psi_s, const float* fmax);

/* number of land cells found by h9_set_soil; cell_xy (optional, n entries)
 * receives (y-1)*lon_c+(x-1) of each land cell in compact order. */
int64_t h9_num_land(const h9_ctx* ctx);
int h9_get_land_index(const h9_ctx* ctx, int32_t* cell_xy);

/* State as INIT.f90:707-811 leaves it (or as a previous h9_get_state returned it):
 * h2osoi_liq (8,lon_c,lat_c); zwt, wa, lai, lai_litter (lon_c,lat_c);
 * plant_mass, plant_foliage_mass, plant_length, rdepth (1,lon_c,lat_c);
 * rootr_col (9,lon_c,lat_c); nplants (lon_c,lat_c) int32;
 * smp (8,lon_c,lat_c) or NULL => zeros (the reference leaves smp uninitialised,
 * INIT.f90:109; see DESIGN.md "smp").  Only land cells are read/written. */
int h9_set_state(h9_ctx* ctx, const float* h2osoi_liq, const float* zwt, const float* wa,
                 const float* lai, const float* lai_litter, const float* plant_mass,
                 const float* plant_foliage_mass, const float* plant_length, const float* rdepth,
                 const float* rootr_col, const int32_t* nplants, const float* smp);
int h9_get_state(h9_ctx* ctx, float* h2osoi_liq, float* zwt, float* wa, float* lai,
                 float* lai_litter, float* plant_mass, float* plant_foliage_mass,
                 float* plant_length, float* rdepth, float* rootr_col, int32_t* nplants, float* smp);

/* ---- the hot path ------------------------------------------------------ */

/* Replaces HYBRID9.f90:120-295 for `ndays` consecutive days: for every land
 * cell, per day: forcing derivation (:156-190), NISURF x HYDROLOGY (:193-211),
 * GROW (:217), daily/annual accumulators (:235-254).
 * year_index_of_day[d] = iY (1-based year within the run, HYBRID9.f90:269) of day d;
 * tas,rlds,rsds,huss,ps,pr,rhs: (lon_c,lat_c,ndays) exactly as READ_PGF leaves them.
 * Host pointers may be pageable or page-locked (h9_host_alloc).  Page-locked arrays
 * are not copied at all: the stepping kernel reads each day's values for its land
 * cells in place over PCIe, one day ahead of use (only ~26 % of a global grid is
 * land).  Pageable arrays are copied tile by tile through a pinned ring.
 * Returns 0, or >0 (fault bits) if any cell faulted -- the caller then reads
 * h9_get_fault and STOPs like the reference. */
int h9_run_days(h9_ctx* ctx, int ndays, const int32_t* year_index_of_day, const float* tas,
                const float* rlds, const float* rsds, const float* huss, const float* ps,
                const float* pr, const float* rhs);

/* Same, with forcing already device-resident in compact form: 7 arrays
 * [ndays][ncell_stride] (day-major, land cells in compact order).  Used by
 * bench.py's kernel-only `value` and by multi-decade runs that keep a forcing
 * ring on the device. Order of the 7 planes: tas,rlds,rsds,huss,ps,pr,rhs. */
int h9_run_days_device(h9_ctx* ctx, int ndays, const int32_t* year_index_of_day,
                       const float* d_forcing, size_t day_stride, size_t plane_stride);

/* Pack (lon_c,lat_c,ndays) host forcing into the compact device layout that
 * h9_run_days_device consumes (land compaction + transpose on the GPU). The
 * returned device pointer is owned by the ctx and valid until the next pack. */
int h9_pack_forcing(h9_ctx* ctx, int ndays, const float* tas, const float* rlds,
                    const float* rsds, const float* huss, const float* ps, const float* pr,
                    const float* rhs, const float** d_forcing, size_t* day_stride,
                    size_t* plane_stride);

/* Annual means of year iyr (1-based) as HYBRID9.f90:263-291 computes them.
 * axy_npp, axy_plant_mass, axy_rnf, axy_evap, axy_theta_total: (lon_c,lat_c);
 * axy_theta: (8,lon_c,lat_c).  Only land cells are written, so the NaN/0 fill
 * of INIT.f90:402-414 survives elsewhere.  Any pointer may be NULL. */
int h9_get_annual(h9_ctx* ctx, int iyr, float* axy_npp, float* axy_plant_mass, float* axy_rnf,
                  float* axy_evap, float* axy_theta_total, float* axy_theta);

/* Compact device-side view of the same means for collectives (no host copy):
 * 13 planes [13][ncell_stride] in the order npp, plant_mass, rnf, evap,
 * theta_total, theta(1..8); plus 8 FP64 budget partial sums over the cells of
 * this ctx: sum of soil water (current h2osoi_liq), sum of wa, sum of annual
 * rnf, sum of annual npp, sum of annual mean plant_mass, n_cells, sum of annual
 * mean total soil water, n_faulted.  Pointers are device pointers valid until
 * the next call; the work is stream-ordered on h9_stream().  The budget is new
 * (the reference has no global reduction); it is FP64 because its summation
 * order depends on the shard count. */
int h9_annual_device(h9_ctx* ctx, int iyr, const float** d_means, size_t* plane_stride,
                     const double** d_budget);

int h9_get_fault(h9_ctx* ctx, h9_fault* out);
int h9_clear_fault(h9_ctx* ctx);

/* CUDA stream (cudaStream_t) the ctx launches on, and a full synchronize. */
void* h9_stream(h9_ctx* ctx);
int h9_synchronize(h9_ctx* ctx);

/* Pinned host memory for forcing arrays (Fortran side: C_F_POINTER on the result). */
void* h9_host_alloc(size_t bytes);
void h9_host_free(void* p);

/* counters for the bench: kernels launched / bytes copied by this ctx so far */
int64_t h9_launch_count(const h9_ctx* ctx);
int64_t h9_h2d_bytes(const h9_ctx* ctx);
int64_t h9_d2h_bytes(const h9_ctx* ctx);
/* device time (ms) spent in the time-stepping kernel since the last reset, by
 * CUDA events on the launching stream; and reset. */
double h9_step_kernel_ms(h9_ctx* ctx);
int h9_reset_counters(h9_ctx* ctx);

/* ---- fine-grained entries (1:1 parity targets for the two routines) ---- */

/* One HYDROLOGY sub-step (HYDROLOGY.f90:141-1283) for all land cells with the
 * given day's forcing, (lon_c,lat_c) each.  Optional outputs (lon_c,lat_c)
 * unless noted: theta (8,lon_c,lat_c), qflx_tran_veg_col, qflx_evap_grnd,
 * rnf_inc ((qflx_surf+rsub_top)*dt), w_imbalance (w1-w0), jwt (int32, the
 * value after Drainage). */
int h9_hydrology_step(h9_ctx* ctx, const float* tas, const float* rlds, const float* rsds,
                      const float* huss, const float* ps, const float* pr, const float* rhs,
                      float* theta, float* qflx_tran_veg_col, float* qflx_evap_grnd,
                      float* rnf_inc, float* w_imbalance, int32_t* jwt);

/* One GROW call (GROW.f90:55-201) for all land cells; tas (lon_c,lat_c).
 * Optional outputs (lon_c,lat_c): npp, w_i, fT. */
int h9_grow_day(h9_ctx* ctx, const float* tas, float* npp, float* w_i, float* fT);

/* INIT's soil pre-processing for ONE layer (INIT.f90:573-633): the four BNU fields at 30
 * arc-seconds, each (lon_c*60, lat_c*60) with x fastest, are block-averaged 60x60 to the
 * half-degree block over the cells whose theta_s_in >= 0 (sums in the reference's order:
 * x1 outer, y1 inner; bit-exact), then converted: theta_s = mean/1e3, hksat = 10*mean/86400,
 * lambda = max(mean/1e3, 1e-8), bsw = 1/lambda, psi_s = 10*mean.  Results go to element
 * `layer` (1..8) of the (8,lon_c,lat_c) arrays.  Host pointers, pageable or pinned; the
 * fine grids are staged through the device in latitude bands.  Only needs h9_create. */
int h9_regrid_soil_layer(h9_ctx* ctx, int lon_c, int lat_c, int layer, const float* theta_s_in,
                         const float* k_s_in, const float* lambda_in, const float* psi_s_in,
                         float* theta_s, float* hksat, float* bsw, float* psi_s);

/* Balanced contiguous latitude bands for `nranks` ranks: rows are assigned so
 * that every band holds about the same number of land cells (the reference's
 * INIT.f90:271-283 uses equal-area squares instead).  lat_s[r] is 1-based,
 * lat_c[r] the row count.  Pure host function, no ctx needed. */
int h9_partition_lat_bands(int lon_c, int lat_c, const int32_t* soil_tex, const float* theta_s,
                           int nranks, int32_t* lat_s, int32_t* lat_count, int64_t* n_land);

/* ---- multi-GPU: one ctx per GPU (one per MPI rank), NCCL on the ctx's stream ---------
 * The reference's ranks exchange nothing during time stepping (HYDROLOGY.f90 indexes only
 * the current (x,y)) and meet again in the collective netCDF writers
 * (WRITE_NET_CDF_3DR.f90:93-94,236-257).  These entries give the host the same meeting point
 * on the device: rank 0 creates an id and broadcasts its H9_COMM_ID_BYTES bytes with the
 * host's own transport (MPI_Bcast after INIT.f90:26-38), every rank calls h9_comm_init, and
 * once per simulated year h9_annual_collective enqueues on h9_stream(), right behind the
 * stepping kernel, with no host synchronisation and persistent buffers: the budget kernel, an
 * FP64 all-reduce of the 8 budget sums of h9_annual_device in the year's own slot, and a ragged
 * all-gather of every rank's 13 annual-mean planes.  (Env H9_COMM_OVERLAP=1 moves the two NCCL
 * operations to a second stream behind an event; measured slower where ranks finish a year at
 * different times, see h9_api.cu.)  libnccl is loaded on first use
 * (dlopen "libnccl.so.2", or the path in env H9_NCCL_LIB); errors return H9_ERR_NCCL. */
#define H9_COMM_ID_BYTES 128
int h9_comm_unique_id(void* id /* H9_COMM_ID_BYTES */);
int h9_comm_init(h9_ctx* ctx, int nranks, int rank, const void* id);
int h9_comm_destroy(h9_ctx* ctx);
/* land cells of every rank's block, n_land[nranks] (collective on first use) */
int h9_comm_land_counts(h9_ctx* ctx, int64_t* n_land);
/* collective: every rank calls it for the same iyr; asynchronous on h9_stream() */
int h9_annual_collective(h9_ctx* ctx, int iyr);
/* results of the last h9_annual_collective (synchronise the stream): rank r's compact
 * planes [13][n_land[r]] in the order of h9_annual_device, and the all-reduced budget of
 * year iyr, budget[8] */
int h9_get_gathered_annual(h9_ctx* ctx, int r, float* planes);
int h9_get_budget(h9_ctx* ctx, int iyr, double* budget);
/* device views for stream-ordered consumers: all ranks' planes back to back, rank r's block
 * [13][stride_r] (stride_r = n_land[r] rounded up to 128); budget slots [nyr][8].  Work enqueued
 * on h9_stream() after this call is ordered behind the last h9_annual_collective */
int h9_gathered_device(h9_ctx* ctx, const float** d_planes, const double** d_budget_years);

#ifdef __cplusplus
}
#endif
#endif /* H9GPU_H */
