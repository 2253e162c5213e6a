/*
 * ref_harness.cpp -- C ABI around the reference's own HYDROLOGY / GROW / loop nest as
 * translated from its Fortran by oracle/f2cpp.py (oracle/_ref/h9_ref_gen.h).
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/h9_oracle.h): libh9ref.so is the checker that pins
 * the hand-written oracle and, through it, the GPU path; the product never loads it.
 *
 * This file holds NO physics.  It plays the parts of PROGRAM H9 / SUBROUTINE INIT that are
 * MPI, netCDF or driver.txt I/O in the reference: it sets the module variables those parts
 * would set (lon_c, lat_c, NISURF, NYR, zi, the soil fields, the forcing arrays), calls the
 * translated units in the reference's order, and copies module arrays in and out.  Every
 * number comes out of translated reference statements.
 *
 *   h9r_configure        INIT.f90:56-135 (scratch), :154 (sla), :214 (dt), :252-263 (dz, zc),
 *                        :301-395 (grid arrays), :402-414 (fills), :844-859 (time_BOY)
 *   h9r_init_state       INIT.f90:711-811
 *   h9r_run_days         HYBRID9.f90:120-295, the loop nest verbatim, on a calendar table
 *                        built from the caller's year index (so partial years can be run);
 *                        per_cell_smp=1 runs the same body cell by cell (HYBRID9.f90:126-292)
 *                        and gives every cell its own `smp` (see DESIGN.md on the smp leak)
 *   h9r_run_decade       HYBRID9.f90:103-113 + :120-295 on the reference's own calendar
 *   h9r_hydrology_step   HYBRID9.f90:156-189 + one CALL HYDROLOGY for every land cell
 *   h9r_grow_day         one CALL GROW for every land cell
 *   h9r_regrid_soil_layer INIT.f90:360-370 + :575-632, the 30 arc-second -> half-degree block means
 *
 * STOP (HYDROLOGY.f90:811,824,1071,1273) throws; the harness records which STOP, the cell,
 * the day and the sub-step, and returns the corresponding H9_FAULT_* bit.
 */
#include <cstdint>
#include <cstring>
#include <vector>

#include "_ref/h9_ref_gen.h"

struct H9Stop {
  int ordinal, line;
};
void H9Ref::f2c_stop(int ordinal, int line) { throw H9Stop{ordinal, line}; }

namespace {
const uint32_t kFaultOfStop[5] = {0u, 1u, 2u, 4u, 8u}; /* order of the STOPs in HYDROLOGY.f90 */
}

struct h9r_ctx {
  H9Ref* r = nullptr;
  int lon_c = 0, lat_c = 0, nisurf = 0, nyr = 0;
  bool have_soil = false;
  std::vector<float> smp_cell; /* (8,lon_c,lat_c): per-cell smp for per_cell_smp=1 */
  std::vector<int32_t> land;   /* 0-based (y*lon_c + x) of land cells, reference order */
  /* first (and only: the reference stops) fault */
  uint32_t fault_code = 0;
  int fault_x = 0, fault_y = 0, fault_day = 0, fault_sub = 0, fault_line = 0;
  float fault_imb = 0.0f;
  ~h9r_ctx() { delete r; }
};

namespace {

void record_stop(h9r_ctx* c, const H9Stop& s, int substep) {
  H9Ref& r = *c->r;
  c->fault_code = (s.ordinal >= 1 && s.ordinal <= 4) ? kFaultOfStop[s.ordinal] : 0x80000000u;
  c->fault_x = r.x_;
  c->fault_y = r.y_;
  c->fault_day = r.it_;
  c->fault_sub = substep;
  c->fault_line = s.line;
  c->fault_imb = r.hydrology__w1_ - r.hydrology__w0_;
}

void ensure_forcing(h9r_ctx* c, int ndays) {
  H9Ref& r = *c->r;
  if (r.tas_.allocated() && r.ntimes_ == ndays) return;
  if (r.tas_.allocated()) { /* READ_PGF.f90:111-112 region: the reference DEALLOCATEs per decade */
    r.tas_.dealloc();
    r.rlds_.dealloc();
    r.rsds_.dealloc();
    r.huss_.dealloc();
    r.ps_.dealloc();
    r.pr_.dealloc();
    r.rhs_.dealloc();
  }
  r.ntimes_ = ndays;
  r.alloc_forcing();
}

void load_forcing(h9r_ctx* c, int ndays, const float* tas, const float* rlds, const float* rsds,
                  const float* huss, const float* ps, const float* pr, const float* rhs) {
  ensure_forcing(c, ndays);
  H9Ref& r = *c->r;
  const size_t n = (size_t)c->lon_c * c->lat_c * ndays * sizeof(float);
  std::memcpy(r.tas_.data(), tas, n);
  std::memcpy(r.rlds_.data(), rlds, n);
  std::memcpy(r.rsds_.data(), rsds, n);
  std::memcpy(r.huss_.data(), huss, n);
  std::memcpy(r.ps_.data(), ps, n);
  std::memcpy(r.pr_.data(), pr, n);
  std::memcpy(r.rhs_.data(), rhs, n);
}

void smp_in(h9r_ctx* c, int32_t ci) {
  for (int I = 1; I <= 8; ++I) c->r->smp_(I) = c->smp_cell[8 * (size_t)ci + I - 1];
}
void smp_out(h9r_ctx* c, int32_t ci) {
  for (int I = 1; I <= 8; ++I) c->smp_cell[8 * (size_t)ci + I - 1] = c->r->smp_(I);
}

template <class T, int R>
void copy_in(FArr<T, R>& a, const T* src) {
  if (src) std::memcpy(a.data(), src, (size_t)a.size() * sizeof(T));
}
template <class T, int R>
void copy_out(FArr<T, R>& a, T* dst) {
  if (dst) std::memcpy(dst, a.data(), (size_t)a.size() * sizeof(T));
}

} /* namespace */

extern "C" {

int h9r_create(h9r_ctx** ctx) {
  if (!ctx) return -1;
  *ctx = new h9r_ctx();
  return 0;
}

int h9r_destroy(h9r_ctx* ctx) {
  delete ctx;
  return 0;
}

int h9r_configure(h9r_ctx* c, int lon_c, int lat_c, int nisurf, const float zi[10], int nyr) {
  if (!c || lon_c < 1 || lat_c < 1 || nisurf < 1 || !zi || nyr < 1) return -1;
  delete c->r;
  c->r = new H9Ref();
  H9Ref& r = *c->r;
  c->lon_c = lon_c, c->lat_c = lat_c, c->nisurf = nisurf, c->nyr = nyr;
  /* what MPI start-up, driver.txt and the block decomposition set (INIT.f90:26-50,172-200,271-296) */
  r.my_id_ = 0;
  r.num_procs_ = 1;
  r.lon_c_ = lon_c;
  r.lat_c_ = lat_c;
  r.lon_s_ = 1;
  r.lat_s_ = 1;
  r.nisurf_ = nisurf;
  r.nyr_ = nyr;
  r.pgf_ = true;
  r.interactive_ = false;
  r.lclim_ = false;
  r.idec_start_ = 1;
  r.idec_end_ = 1;
  r.init_alloc_scratch();                            /* INIT.f90:56-135 */
  for (int I = 0; I <= 9; ++I) r.zi_(I) = zi[I];     /* driver.txt, INIT.f90:197-199 */
  r.init_gpt();                                      /* :154 */
  r.init_dt();                                       /* :214 */
  r.init_layers();                                   /* :252-263 */
  r.init_alloc_grid_a();                             /* :301-355 */
  r.init_alloc_grid_b();                             /* :374-395 */
  r.init_fills();                                    /* :402-414 */
  r.init_time_boy();                                 /* :844-859 */
  r.nlayers_ = r.nsoil_layers_max_;
  c->smp_cell.assign((size_t)8 * lon_c * lat_c, 0.0f);
  c->land.clear();
  c->have_soil = false;
  c->fault_code = 0;
  return 0;
}

/* the soil fields as INIT.f90:604-633 leaves them (netCDF read + regridding replaced by the
 * caller's arrays); lambda = 1/bsw inverts :628, theta_ma_s = 0.1 is :626 */
int h9r_set_soil(h9r_ctx* c, const int32_t* soil_tex, const float* theta_s, const float* hksat,
                 const float* bsw, const float* psi_s, const float* fmax) {
  if (!c || !c->r || !soil_tex || !theta_s || !hksat || !bsw || !psi_s || !fmax) return -1;
  H9Ref& r = *c->r;
  copy_in(r.soil_tex_, soil_tex);
  copy_in(r.theta_s_, theta_s);
  copy_in(r.hksat_, hksat);
  copy_in(r.bsw_, bsw);
  copy_in(r.psi_s_, psi_s);
  copy_in(r.fmax_, fmax);
  for (long k = 0; k < r.bsw_.size(); ++k) {
    r.lambda_.data()[k] = 1.0f / r.bsw_.data()[k];
    r.theta_ma_s_.data()[k] = 0.1f;
  }
  /* land list for the per-cell entry points; the verbatim loop nest applies the reference's
   * own test (HYBRID9.f90:122-123) and tests/test_ref_vs_oracle.py checks both agree */
  c->land.clear();
  for (int y = 1; y <= c->lat_c; ++y)
    for (int x = 1; x <= c->lon_c; ++x) {
      float s = 0.0f;
      for (int I = 1; I <= 8; ++I) s = s + r.theta_s_(I, x, y);
      if (r.soil_tex_(x, y) > 0 && r.soil_tex_(x, y) != 13 && s > r.trunc_)
        c->land.push_back((int32_t)((y - 1) * c->lon_c + (x - 1)));
    }
  c->have_soil = true;
  return 0;
}

int64_t h9r_num_land(const h9r_ctx* c) { return c ? (int64_t)c->land.size() : -1; }

int h9r_get_land_index(const h9r_ctx* c, int32_t* cell_xy) {
  if (!c || !cell_xy) return -1;
  std::memcpy(cell_xy, c->land.data(), c->land.size() * sizeof(int32_t));
  return 0;
}

int h9r_init_state(h9r_ctx* c) {
  if (!c || !c->have_soil) return -1;
  c->r->init_state(); /* INIT.f90:711-811 */
  std::fill(c->smp_cell.begin(), c->smp_cell.end(), 0.0f);
  for (int I = 1; I <= 8; ++I) c->r->smp_(I) = 0.0f;
  return 0;
}

int h9r_set_state(h9r_ctx* c, const float* h2osoi_liq, const float* zwt, const float* wa,
                  const float* lai, const float* lai_litter, const float* plant_mass,
                  const float* plant_foliage_mass, const float* plant_length, const float* rdepth,
                  const float* rootr_col, const int32_t* nplants, const float* smp) {
  if (!c || !c->have_soil) return -1;
  H9Ref& r = *c->r;
  copy_in(r.h2osoi_liq_, h2osoi_liq);
  copy_in(r.zwt_, zwt);
  copy_in(r.wa_, wa);
  copy_in(r.lai_, lai);
  copy_in(r.lai_litter_, lai_litter);
  copy_in(r.plant_mass_, plant_mass);
  copy_in(r.plant_foliage_mass_, plant_foliage_mass);
  copy_in(r.plant_length_, plant_length);
  copy_in(r.rdepth_, rdepth);
  copy_in(r.rootr_col_, rootr_col);
  copy_in(r.nplants_, nplants);
  if (smp)
    std::memcpy(c->smp_cell.data(), smp, c->smp_cell.size() * sizeof(float));
  else
    std::fill(c->smp_cell.begin(), c->smp_cell.end(), 0.0f);
  for (int I = 1; I <= 8; ++I) r.smp_(I) = 0.0f;
  return 0;
}

int h9r_get_state(h9r_ctx* c, float* h2osoi_liq, float* zwt, float* wa, float* lai,
                  float* lai_litter, float* plant_mass, float* plant_foliage_mass,
                  float* plant_length, float* rdepth, float* rootr_col, int32_t* nplants,
                  float* smp, float* smp_shared /* [8]: the module's one scratch vector */) {
  if (!c || !c->have_soil) return -1;
  H9Ref& r = *c->r;
  copy_out(r.h2osoi_liq_, h2osoi_liq);
  copy_out(r.zwt_, zwt);
  copy_out(r.wa_, wa);
  copy_out(r.lai_, lai);
  copy_out(r.lai_litter_, lai_litter);
  copy_out(r.plant_mass_, plant_mass);
  copy_out(r.plant_foliage_mass_, plant_foliage_mass);
  copy_out(r.plant_length_, plant_length);
  copy_out(r.rdepth_, rdepth);
  copy_out(r.rootr_col_, rootr_col);
  copy_out(r.nplants_, nplants);
  if (smp) std::memcpy(smp, c->smp_cell.data(), c->smp_cell.size() * sizeof(float));
  if (smp_shared)
    for (int I = 1; I <= 8; ++I) smp_shared[I - 1] = r.smp_(I);
  return 0;
}

/* Runs HYBRID9.f90:120-295 over `ndays` days.  year_index_of_day must be non-decreasing in
 * steps of at most 1; year index k is calendar year 1900+k of the table (iDEC_start = 1, so
 * iY = k, HYBRID9.f90:265).  Returns 0 or the H9_FAULT_* bit of the STOP that ended the run. */
int h9r_run_days(h9r_ctx* c, int ndays, const int32_t* year_index_of_day, const float* tas,
                 const float* rlds, const float* rsds, const float* huss, const float* ps,
                 const float* pr, const float* rhs, int per_cell_smp) {
  if (!c || !c->have_soil || ndays < 1 || !year_index_of_day) return -1;
  if (!tas || !rlds || !rsds || !huss || !ps || !pr || !rhs) return -1;
  H9Ref& r = *c->r;
  const int iy0 = year_index_of_day[0], iy1 = year_index_of_day[ndays - 1];
  if (iy0 < 1 || iy1 > c->nyr) return -1;
  for (int d = 1; d < ndays; ++d) {
    const int s = year_index_of_day[d] - year_index_of_day[d - 1];
    if (s < 0 || s > 1) return -1;
  }
  load_forcing(c, ndays, tas, rlds, rsds, huss, ps, pr, rhs);
  /* calendar table for these days: time_BOY(jyear-1859) = first day number of year jyear */
  std::vector<int> saved((size_t)r.time_boy_.size());
  std::memcpy(saved.data(), r.time_boy_.data(), saved.size() * sizeof(int));
  r.idec_start_ = 1;
  r.syr_ = 1900 + iy0;
  r.eyr_ = 1900 + iy1;
  int t = r.time_boy_(r.syr_ - 1859);
  for (int iy = iy0, d = 0; iy <= iy1; ++iy) {
    int cnt = 0;
    while (d < ndays && year_index_of_day[d] == iy) ++d, ++cnt;
    r.time_boy_(1900 + iy - 1859) = t;
    t += cnt;
    r.time_boy_(1900 + iy + 1 - 1859) = t;
  }
  int rc = 0;
  c->fault_code = 0;
  try {
    if (!per_cell_smp) {
      r.decade_loop(); /* HYBRID9.f90:120-295 verbatim */
    } else {
      for (size_t k = 0; k < c->land.size(); ++k) {
        const int32_t ci = c->land[k];
        r.y_ = ci / c->lon_c + 1;
        r.x_ = ci % c->lon_c + 1;
        smp_in(c, ci);
        r.cell_years(); /* HYBRID9.f90:126-292 */
        smp_out(c, ci);
      }
    }
  } catch (const H9Stop& s) {
    record_stop(c, s, r.ns_);
    rc = (int)c->fault_code;
  }
  std::memcpy(r.time_boy_.data(), saved.data(), saved.size() * sizeof(int));
  return rc;
}

/* One decade on the reference's own calendar: iDEC as in HYBRID9.f90:93; ndays must equal
 * time_BOY(eyr+1-1859) - time_BOY(syr-1859) (3652/3653 days; 731 for iDEC = 12). */
int h9r_run_decade(h9r_ctx* c, int idec_start, int idec, int ndays, const float* tas,
                   const float* rlds, const float* rsds, const float* huss, const float* ps,
                   const float* pr, const float* rhs) {
  if (!c || !c->have_soil) return -1;
  H9Ref& r = *c->r;
  r.idec_start_ = idec_start;
  r.idec_ = idec;
  r.decade_years(); /* HYBRID9.f90:103-113 */
  const int nt = r.time_boy_(r.eyr_ + 1 - 1859) - r.time_boy_(r.syr_ - 1859);
  if (nt != ndays) return -2;
  if (r.eyr_ - ((idec_start - 1) * 10 + 1901) + 1 > c->nyr) return -3;
  load_forcing(c, ndays, tas, rlds, rsds, huss, ps, pr, rhs);
  c->fault_code = 0;
  try {
    r.decade_loop();
  } catch (const H9Stop& s) {
    record_stop(c, s, r.ns_);
    return (int)c->fault_code;
  }
  return 0;
}

int h9r_get_annual(h9r_ctx* c, int iyr, float* axy_npp, float* axy_plant_mass, float* axy_rnf,
                   float* axy_evap, float* axy_theta_total, float* axy_theta) {
  if (!c || !c->have_soil || iyr < 1 || iyr > c->nyr) return -1;
  H9Ref& r = *c->r;
  const size_t n = (size_t)c->lon_c * c->lat_c, o = (size_t)(iyr - 1) * n;
  if (axy_npp) std::memcpy(axy_npp, r.axy_npp_.data() + o, n * 4);
  if (axy_plant_mass) std::memcpy(axy_plant_mass, r.axy_plant_mass_.data() + o, n * 4);
  if (axy_rnf) std::memcpy(axy_rnf, r.axy_rnf_.data() + o, n * 4);
  if (axy_evap) std::memcpy(axy_evap, r.axy_evap_.data() + o, n * 4);
  if (axy_theta_total) std::memcpy(axy_theta_total, r.axy_theta_total_.data() + o, n * 4);
  if (axy_theta) std::memcpy(axy_theta, r.axy_theta_.data() + 8 * o, 8 * n * 4);
  return 0;
}

/* the 7 forcing means of HYBRID9.f90:277-283, order tas rlds rsds huss ps pr rhs */
int h9r_get_annual_forcing(h9r_ctx* c, int iyr, float* out7) {
  if (!c || !c->have_soil || iyr < 1 || iyr > c->nyr || !out7) return -1;
  H9Ref& r = *c->r;
  const size_t n = (size_t)c->lon_c * c->lat_c, o = (size_t)(iyr - 1) * n;
  const float* src[7] = {r.axy_tas_.data(), r.axy_rlds_.data(), r.axy_rsds_.data(),
                         r.axy_huss_.data(), r.axy_ps_.data(),  r.axy_pr_.data(),
                         r.axy_rhs_.data()};
  for (int f = 0; f < 7; ++f) std::memcpy(out7 + f * n, src[f] + o, n * 4);
  return 0;
}

#define H9R_NDIAG 16
/* One CALL HYDROLOGY for every land cell with one day of forcing (lon_c,lat_c).
 * diag (nullable): (H9R_NDIAG,lon_c,lat_c) locals of the call, in the order
 * qflx_surf rsub_top qflx_rsub_sat qflx_infl qcharge fsat beta rsc w0 w1 rous zwtmm
 * desatdT gamma rho Rnets.  Returns the OR of the fault bits (a STOP only ends that cell). */
int h9r_hydrology_step(h9r_ctx* c, const float* tas, const float* rlds, const float* rsds,
                       const float* huss, const float* ps, const float* pr, const float* rhs,
                       float* theta, float* qflx_tran_veg_col, float* qflx_evap_grnd,
                       float* rnf_inc, float* w_imbalance, int32_t* jwt, float* diag,
                       int per_cell_smp) {
  if (!c || !c->have_soil) return -1;
  if (!tas || !rlds || !rsds || !huss || !ps || !pr || !rhs) return -1;
  H9Ref& r = *c->r;
  load_forcing(c, 1, tas, rlds, rsds, huss, ps, pr, rhs);
  r.syr_ = 1901;
  r.jyear_ = 1901;
  r.itime_ = r.time_boy_(r.syr_ - 1859); /* => iT = 1, DOY = 1 (HYBRID9.f90:156-157) */
  r.nlayers_ = r.nsoil_layers_max_;
  uint32_t any = 0;
  c->fault_code = 0;
  for (size_t k = 0; k < c->land.size(); ++k) {
    const int32_t ci = c->land[k];
    r.y_ = ci / c->lon_c + 1;
    r.x_ = ci % c->lon_c + 1;
    if (per_cell_smp) smp_in(c, ci);
    r.day_derive(); /* HYBRID9.f90:156-189 */
    r.rnf_sum_ = 0.0f;
    r.ns_ = 1;
    try {
      r.hydrology();
    } catch (const H9Stop& s) {
      if (!c->fault_code) record_stop(c, s, 1);
      any |= (s.ordinal >= 1 && s.ordinal <= 4) ? kFaultOfStop[s.ordinal] : 0x80000000u;
    }
    if (per_cell_smp) smp_out(c, ci);
    if (theta)
      for (int I = 1; I <= 8; ++I) theta[8 * (size_t)ci + I - 1] = r.theta_(I);
    if (qflx_tran_veg_col) qflx_tran_veg_col[ci] = r.qflx_tran_veg_col_;
    if (qflx_evap_grnd) qflx_evap_grnd[ci] = r.qflx_evap_grnd_;
    if (rnf_inc) rnf_inc[ci] = r.rnf_sum_;
    if (w_imbalance) w_imbalance[ci] = r.hydrology__w1_ - r.hydrology__w0_;
    if (jwt) jwt[ci] = r.jwt_;
    if (diag) {
      float* d = diag + (size_t)H9R_NDIAG * ci;
      d[0] = r.hydrology__qflx_surf_;
      d[1] = r.hydrology__rsub_top_;
      d[2] = r.hydrology__qflx_rsub_sat_;
      d[3] = r.hydrology__qflx_infl_;
      d[4] = r.hydrology__qcharge_;
      d[5] = r.hydrology__fsat_;
      d[6] = r.hydrology__beta_;
      d[7] = r.hydrology__rsc_;
      d[8] = r.hydrology__w0_;
      d[9] = r.hydrology__w1_;
      d[10] = r.hydrology__rous_;
      d[11] = r.hydrology__zwtmm_;
      d[12] = r.hydrology__desatdt_;
      d[13] = r.hydrology__gamma_;
      d[14] = r.hydrology__rho_;
      d[15] = r.hydrology__rnets_;
    }
  }
  return (int)any;
}

/* One CALL GROW for every land cell; tas is one day (lon_c,lat_c) */
int h9r_grow_day(h9r_ctx* c, const float* tas, float* npp, float* w_i, float* fT,
                 int per_cell_smp) {
  if (!c || !c->have_soil || !tas) return -1;
  H9Ref& r = *c->r;
  ensure_forcing(c, 1);
  std::memcpy(r.tas_.data(), tas, (size_t)c->lon_c * c->lat_c * sizeof(float));
  r.it_ = 1;
  r.nlayers_ = r.nsoil_layers_max_;
  for (size_t k = 0; k < c->land.size(); ++k) {
    const int32_t ci = c->land[k];
    r.y_ = ci / c->lon_c + 1;
    r.x_ = ci % c->lon_c + 1;
    if (per_cell_smp) smp_in(c, ci);
    r.grow();
    if (npp) npp[ci] = r.npp_;
    if (w_i) w_i[ci] = r.w_i_;
    if (fT) fT[ci] = r.ft_;
  }
  return 0;
}

int h9r_get_fault(h9r_ctx* c, uint32_t* code, int32_t* x, int32_t* y, int32_t* day,
                  int32_t* substep, float* imbalance, int32_t* line) {
  if (!c) return -1;
  if (code) *code = c->fault_code;
  if (x) *x = c->fault_code ? c->fault_x : 0;
  if (y) *y = c->fault_code ? c->fault_y : 0;
  if (day) *day = c->fault_code ? c->fault_day : 0;
  if (substep) *substep = c->fault_code ? c->fault_sub : 0;
  if (imbalance) *imbalance = c->fault_code ? c->fault_imb : 0.0f;
  if (line) *line = c->fault_code ? c->fault_line : 0;
  return 0;
}

int h9r_get_geometry(h9r_ctx* c, float dz[10], float zc[10], float* dt) {
  if (!c || !c->r) return -1;
  H9Ref& r = *c->r;
  if (dz) dz[0] = 0.0f;
  if (zc) zc[0] = 0.0f;
  for (int I = 1; I <= 9; ++I) {
    if (dz) dz[I] = r.dz_(I);
    if (zc) zc[I] = r.zc_(I);
  }
  if (dt) *dt = r.dt_;
  return 0;
}

/* time_BOY(year-1859) as INIT.f90:844-859 computed it */
int h9r_time_boy(h9r_ctx* c, int year) {
  if (!c || !c->r || year < 1860 || year > 2300) return -1;
  return c->r->time_boy_(year - 1859);
}

/* INIT.f90:575-632 for one soil layer.  Inputs are the four
 * 30-arc-second tiles (lon_c*60, lat_c*60) as NF90_GET_VAR leaves them (:541-568); outputs are
 * (8,lon_c,lat_c) with only `layer` written, like h9o_regrid_soil_layer. */
int h9r_regrid_soil_layer(int lon_c, int lat_c, int layer, const float* theta_s_in,
                          const float* k_s_in, const float* lambda_in, const float* psi_s_in,
                          float* theta_s, float* hksat, float* bsw, float* psi_s) {
  if (lon_c < 1 || lat_c < 1 || layer < 1 || layer > 8) return -1;
  H9Ref r;
  r.lon_c_ = lon_c;
  r.lat_c_ = lat_c;
  r.init_alloc_l1();     /* INIT.f90:360-370 */
  r.init_alloc_grid_b(); /* :374-395 */
  copy_in(r.theta_s_l1_in_, theta_s_in);
  copy_in(r.k_s_l1_in_, k_s_in);
  copy_in(r.lambda_l1_in_, lambda_in);
  copy_in(r.psi_s_l1_in_, psi_s_in);
  r.i_ = layer; /* the DO I = 1, nsoil_layers_max of :475 */
  r.init_regrid_layer();
  for (int y = 1; y <= lat_c; ++y)
    for (int x = 1; x <= lon_c; ++x) {
      const size_t o = 8 * ((size_t)(y - 1) * lon_c + (x - 1)) + (layer - 1);
      if (theta_s) theta_s[o] = r.theta_s_(layer, x, y);
      if (hksat) hksat[o] = r.hksat_(layer, x, y);
      if (bsw) bsw[o] = r.bsw_(layer, x, y);
      if (psi_s) psi_s[o] = r.psi_s_(layer, x, y);
    }
  return 0;
}

int h9r_ndiag(void) { return H9R_NDIAG; }

} /* extern "C" */
