/*
 * h9_driver.cpp -- C++ host above the C ABI: what PROGRAM H9 (HYBRID9.f90:2-589)
 * does around the hot path when the path itself runs in libh9gpu.so.
 *
 * The reference's host is Fortran; this image has no Fortran compiler, so this
 * program plays its role for the PGF branch (HYBRID9.f90:87-332,492-519):
 *   driver.txt (EXECUTE/driver.txt, read positionally like INIT.f90:181-206)
 *   -> geometry and calendar (INIT.f90:214,252-257,844-859)
 *   -> soil fields + initial state (INIT.f90:707-811) on the host
 *   -> per decade: forcing arrays as READ_PGF leaves them -> h9_run_days
 *      -> STOP with the reference's message on a physics fault
 *      -> h9_get_annual into axy_*(lon_c,lat_c,NYR), forcing means on the host
 *         (HYBRID9.f90:235-241,278-284)
 *   -> raw little-endian dumps of axy_* where the reference calls WRITE_NET_CDF_3DR.
 * netCDF and MPI are out of scope: arrays are exchanged as
 * flat float32/int32 files of the same shapes and memory order.
 *
 * usage: h9_driver <data_dir> [driver.txt] [--exact] [--pageable]
 *   <data_dir>/grid.txt           "lon_c lat_c"
 *   <data_dir>/soil_tex.i32       (lon_c,lat_c)
 *   <data_dir>/{theta_s,hksat,bsw,psi_s}.f32  (8,lon_c,lat_c)
 *   <data_dir>/fmax.f32           (lon_c,lat_c)
 *   <data_dir>/{tas,rlds,rsds,huss,ps,pr,rhs}_dec<NN>.f32  (lon_c,lat_c,NTIMES)
 * outputs: <data_dir>/out_axy_<name>.f32, <data_dir>/out_state_<name>.f32
 */
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/h9gpu.h"

namespace {

struct DriverTxt { /* EXECUTE/driver.txt, INIT.f90:181-206 */
  std::string out_path;
  int NISURF = 48;
  bool PGF = true;
  int iDEC_start = 1, iDEC_end = 1;
  bool INTERACTIVE = false;
  float zi[10] = {0, 45, 91, 166, 289, 493, 829, 1383, 2296, 5000};
};

bool parse_logical(const std::string& s) { return s.find('T') != std::string::npos || s.find('t') != std::string::npos; }

/* list-directed reads: first token of each record, text after '!' ignored */
bool read_driver(const std::string& path, DriverTxt& d) {
  std::ifstream f(path);
  if (!f) return false;
  std::vector<std::string> tok;
  std::string line;
  while (std::getline(f, line)) {
    const size_t bang = line.find('!');
    /* a quoted path may contain '!' only after the closing quote in the reference's file */
    std::string head = line;
    size_t q1 = line.find('\''), q2 = q1 == std::string::npos ? q1 : line.find('\'', q1 + 1);
    if (q2 != std::string::npos) {
      tok.push_back(line.substr(q1 + 1, q2 - q1 - 1));
      continue;
    }
    if (bang != std::string::npos) head = line.substr(0, bang);
    std::istringstream is(head);
    std::string t;
    if (is >> t) tok.push_back(t);
    if (tok.size() >= 26) break; /* 16 scalars + 10 interfaces; the rest of the file is notes */
  }
  if (tok.size() < 26) return false;
  d.out_path = tok[0];
  d.NISURF = atoi(tok[1].c_str());
  d.PGF = parse_logical(tok[2]);
  d.iDEC_start = atoi(tok[3].c_str());
  d.iDEC_end = atoi(tok[4].c_str());
  d.INTERACTIVE = parse_logical(tok[5]);
  for (int i = 0; i < 10; ++i) d.zi[i] = (float)atof(tok[16 + i].c_str());
  return true;
}

int time_BOY(int year) { /* INIT.f90:844-859 */
  int t = 1;
  for (int jyear = 1861; jyear <= year; ++jyear) {
    if ((jyear - 1) % 4 != 0) t += 365;
    else if ((jyear - 1) % 100 != 0) t += 366;
    else if ((jyear - 1) % 400 != 0) t += 365;
    else t += 366;
  }
  return t;
}

template <class T>
bool read_raw(const std::string& path, T* dst, size_t n) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) {
    fprintf(stderr, "cannot open %s\n", path.c_str());
    return false;
  }
  const size_t got = fread(dst, sizeof(T), n, f);
  fclose(f);
  if (got != n) fprintf(stderr, "%s: expected %zu values, got %zu\n", path.c_str(), n, got);
  return got == n;
}

bool write_raw(const std::string& path, const float* src, size_t n) {
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) return false;
  const bool ok = fwrite(src, sizeof(float), n, f) == n;
  fclose(f);
  return ok;
}

} /* namespace */

#define H9CALL(expr)                                                              \
  do {                                                                            \
    int rc_ = (expr);                                                             \
    if (rc_ < 0) {                                                                \
      fprintf(stderr, "%s failed (%d): %s\n", #expr, rc_, h9_last_error(ctx));    \
      return 2;                                                                   \
    }                                                                             \
  } while (0)

int main(int argc, char** argv) {
  if (argc < 2) {
    fprintf(stderr, "usage: %s <data_dir> [driver.txt] [--exact] [--pageable]\n", argv[0]);
    return 1;
  }
  const std::string dir = argv[1];
  std::string driver_path = dir + "/driver.txt";
  bool exact = false, pageable = false;
  for (int i = 2; i < argc; ++i) {
    if (!strcmp(argv[i], "--exact")) exact = true;
    else if (!strcmp(argv[i], "--pageable")) pageable = true;
    else driver_path = argv[i];
  }
  const clock_t start = clock(); /* CPU_TIME (start), INIT.f90:50 */

  DriverTxt drv;
  if (!read_driver(driver_path, drv)) {
    fprintf(stderr, "cannot read %s\n", driver_path.c_str());
    return 1;
  }
  if (!drv.PGF) {
    fprintf(stderr, "only the PGF branch (HYBRID9.f90:87-332) runs on the GPU path\n");
    return 1;
  }
  int lon_c = 0, lat_c = 0;
  {
    std::ifstream g(dir + "/grid.txt");
    if (!(g >> lon_c >> lat_c)) {
      fprintf(stderr, "cannot read %s/grid.txt\n", dir.c_str());
      return 1;
    }
  }
  const size_t ng = (size_t)lon_c * lat_c;
  int NYR; /* INIT.f90:289-293 */
  if (drv.iDEC_end < 12) NYR = (drv.iDEC_end - drv.iDEC_start + 1) * 10;
  else NYR = (drv.iDEC_end - drv.iDEC_start + 1 - 1) * 10 + 2;

  /* soil fields: what INIT.f90:470-680 leaves in SHARED */
  std::vector<int32_t> soil_tex(ng), nplants(ng, 0);
  std::vector<float> theta_s(8 * ng), hksat(8 * ng), bsw(8 * ng), psi_s(8 * ng), Fmax(ng);
  if (!read_raw(dir + "/soil_tex.i32", soil_tex.data(), ng) ||
      !read_raw(dir + "/theta_s.f32", theta_s.data(), 8 * ng) ||
      !read_raw(dir + "/hksat.f32", hksat.data(), 8 * ng) ||
      !read_raw(dir + "/bsw.f32", bsw.data(), 8 * ng) ||
      !read_raw(dir + "/psi_s.f32", psi_s.data(), 8 * ng) ||
      !read_raw(dir + "/fmax.f32", Fmax.data(), ng))
    return 1;

  /* geometry, INIT.f90:252-257 */
  float dz[10] = {0}, zc[10] = {0};
  for (int I = 1; I <= 9; ++I) dz[I] = drv.zi[I] - drv.zi[I - 1];
  for (int I = 1; I <= 9; ++I) zc[I] = drv.zi[I] - dz[I] / 2.0f;
  (void)zc;

  /* initial state, INIT.f90:707-811 */
  std::vector<float> h2osoi_liq(8 * ng, 0.0f), zwt(ng, 0.0f), wa(ng, 0.0f), LAI(ng, 0.0f),
      LAI_litter(ng, 0.0f), plant_mass(ng, 0.0f), plant_foliage_mass(ng, 0.0f),
      plant_length(ng, 0.0f), rdepth(ng, 0.0f), rootr_col(9 * ng, 0.0f), smp(8 * ng, 0.0f);
  const float rhow = 1000.0f, sla = 23.0E-3f, plot_area = 1.0f;
  long nland = 0;
  for (int y = 0; y < lat_c; ++y)
    for (int x = 0; x < lon_c; ++x) {
      const size_t c = (size_t)y * lon_c + x;
      float sum = 0.0f;
      for (int I = 0; I < 8; ++I) sum = sum + theta_s[8 * c + I];
      if (!(soil_tex[c] > 0 && soil_tex[c] != 13 && sum > 1.0E-8f)) continue;
      ++nland;
      for (int I = 1; I <= 8; ++I) h2osoi_liq[8 * c + I - 1] = 0.4f * theta_s[8 * c + I - 1] * dz[I] * rhow / 1000.0f;
      zwt[c] = (drv.zi[8] + 5000.0f) / 1000.0f;
      wa[c] = 4000.0f;
      LAI_litter[c] = 0.001f;
      nplants[c] = 1;
      plant_mass[c] = 1.0f;
      plant_foliage_mass[c] = 0.0435f;
      plant_length[c] = powf(400.0f * plant_mass[c] / 3.142E-3f, 1.0f / 3.0f);
      LAI[c] = 0.0f + plant_foliage_mass[c] * sla / plot_area;
      rdepth[c] = 0.3f * plant_length[c];
      const float decay = expf(logf(0.1f) / (rdepth[c] / 10.0f));
      for (int I = 1; I <= 8; ++I)
        rootr_col[9 * c + I - 1] = rootr_col[9 * c + I - 1] + (1.0f - powf(decay, drv.zi[I] / 10.0f)) -
                                   (1.0f - powf(decay, drv.zi[I - 1] / 10.0f));
    }
  printf("lon_c lat_c %d %d  land cells %ld  NISURF %d  decades %d-%d  NYR %d\n", lon_c, lat_c, nland,
         drv.NISURF, drv.iDEC_start, drv.iDEC_end, NYR);

  /* axy_* with the fills of INIT.f90:402-414 */
  const float nanv = std::nanf("");
  const size_t nyg = (size_t)NYR * ng;
  std::vector<float> axy_npp(nyg, nanv), axy_plant_mass(nyg, nanv), axy_rnf(nyg, nanv), axy_evap(nyg, nanv),
      axy_tas(nyg, nanv), axy_huss(nyg, nanv), axy_ps(nyg, nanv), axy_pr(nyg, nanv), axy_rhs(nyg, nanv),
      axy_theta(8 * nyg, nanv), axy_theta_total(nyg, 0.0f);

  h9_ctx* ctx = nullptr;
  if (h9_create(&ctx, -1) != H9_OK) {
    fprintf(stderr, "h9_create failed: no usable CUDA device (there is no CPU path)\n");
    return 2;
  }
  H9CALL(h9_configure(ctx, lon_c, lat_c, drv.NISURF, drv.zi, NYR));
  H9CALL(h9_set_math(ctx, exact ? H9_MATH_EXACT : H9_MATH_FAST));
  H9CALL(h9_set_soil(ctx, soil_tex.data(), theta_s.data(), hksat.data(), bsw.data(), psi_s.data(), Fmax.data()));
  H9CALL(h9_set_state(ctx, h2osoi_liq.data(), zwt.data(), wa.data(), LAI.data(), LAI_litter.data(),
                      plant_mass.data(), plant_foliage_mass.data(), plant_length.data(), rdepth.data(),
                      rootr_col.data(), nplants.data(), nullptr));

  const char* fname[7] = {"tas", "rlds", "rsds", "huss", "ps", "pr", "rhs"};
  for (int iDEC = drv.iDEC_start; iDEC <= drv.iDEC_end; ++iDEC) { /* HYBRID9.f90:93 */
    const int syr = (iDEC - 1) * 10 + 1901;              /* :103 */
    const int eyr = iDEC < 12 ? syr + 9 : syr + 1;        /* :109-113 */
    const int NTIMES = time_BOY(eyr + 1) - time_BOY(syr);
    /* READ_PGF (:97): seven (lon_c,lat_c,NTIMES) arrays, pinned unless --pageable */
    float* forc[7];
    const size_t nf = ng * (size_t)NTIMES;
    for (int v = 0; v < 7; ++v) {
      forc[v] = pageable ? (float*)malloc(nf * sizeof(float)) : (float*)h9_host_alloc(nf * sizeof(float));
      char nm[64];
      snprintf(nm, sizeof nm, "/%s_dec%02d.f32", fname[v], iDEC);
      if (!forc[v] || !read_raw(dir + nm, forc[v], nf)) return 1;
    }
    std::vector<int32_t> year_of_day(NTIMES);
    for (int jyear = syr; jyear <= eyr; ++jyear) {
      const int iY = jyear - ((drv.iDEC_start - 1) * 10 + 1901) + 1; /* :269 */
      for (int iTIME = time_BOY(jyear); iTIME <= time_BOY(jyear + 1) - 1; ++iTIME)
        year_of_day[iTIME - time_BOY(syr)] = iY;                     /* iT-1, :156 */
    }
    const int rc = h9_run_days(ctx, NTIMES, year_of_day.data(), forc[0], forc[1], forc[2], forc[3], forc[4],
                               forc[5], forc[6]);
    if (rc < 0) {
      fprintf(stderr, "h9_run_days failed (%d): %s\n", rc, h9_last_error(ctx));
      return 2;
    }
    if (rc > 0) { /* the reference's STOPs */
      h9_fault f;
      h9_get_fault(ctx, &f);
      if (f.code & H9_FAULT_WATER_IMBALANCE) {
        printf("\n Problem in HYDROLOGY\n Water imbalance > 0.1 mm  %g\n", f.imbalance);
      } else if (f.code & H9_FAULT_TRIDIAG_PIVOT1) {
        printf(" Problem with tridiagonal 1.\n");
      } else if (f.code & H9_FAULT_TRIDIAG_PIVOT2) {
        printf(" Problem with tridiagonal 2.\n");
      } else {
        printf(" rsub_top_tot is positive in drainage\n HYBRID9 is stopping\n");
      }
      printf(" DiTIME =  %d\n my_id x y  0 %d %d  (sub-step %d, %lld cells faulted)\n", f.day, f.x, f.y,
             f.substep, (long long)f.n_faulted);
      return 3; /* STOP */
    }
    for (int jyear = syr; jyear <= eyr; ++jyear) {
      const int iY = jyear - ((drv.iDEC_start - 1) * 10 + 1901) + 1;
      const size_t o = (size_t)(iY - 1) * ng;
      H9CALL(h9_get_annual(ctx, iY, &axy_npp[o], &axy_plant_mass[o], &axy_rnf[o], &axy_evap[o],
                           &axy_theta_total[o], &axy_theta[8 * o]));
      /* forcing means stay on the host: HYBRID9.f90:235-241,278-284 */
      const int t0 = time_BOY(jyear) - time_BOY(syr), t1 = time_BOY(jyear + 1) - time_BOY(syr);
      const float nt = (float)(t1 - t0);
      const float* src[5] = {forc[0], forc[3], forc[4], forc[5], forc[6]};
      float* dst[5] = {&axy_tas[o], &axy_huss[o], &axy_ps[o], &axy_pr[o], &axy_rhs[o]};
      for (size_t c = 0; c < ng; ++c) {
        if (nplants[c] != 1) continue; /* land cells only, like the loop :120-123 */
        for (int v = 0; v < 5; ++v) {
          float sum = 0.0f;
          for (int t = t0; t < t1; ++t) sum = sum + src[v][(size_t)t * ng + c];
          dst[v][c] = sum / nt;
        }
      }
    }
    for (int v = 0; v < 7; ++v) { /* DEALLOCATE, :322-328 */
      if (pageable) free(forc[v]);
      else h9_host_free(forc[v]);
    }
    printf("decade %d (%d-%d, %d days) done\n", iDEC, syr, eyr, NTIMES);
  }

  /* where the reference calls WRITE_NET_CDF_3DR (HYBRID9.f90:492-519) */
  struct {
    const char* name;
    const std::vector<float>* v;
  } outs[] = {{"npp", &axy_npp}, {"plant_mass", &axy_plant_mass}, {"rnf", &axy_rnf}, {"evap", &axy_evap},
              {"tas", &axy_tas}, {"huss", &axy_huss}, {"ps", &axy_ps}, {"pr", &axy_pr}, {"rhs", &axy_rhs},
              {"theta_total", &axy_theta_total}, {"theta", &axy_theta}};
  for (auto& o : outs) write_raw(dir + "/out_axy_" + o.name + ".f32", o.v->data(), o.v->size());
  H9CALL(h9_get_state(ctx, h2osoi_liq.data(), zwt.data(), wa.data(), LAI.data(), LAI_litter.data(),
                      plant_mass.data(), plant_foliage_mass.data(), plant_length.data(), rdepth.data(),
                      rootr_col.data(), nplants.data(), smp.data()));
  write_raw(dir + "/out_state_h2osoi_liq.f32", h2osoi_liq.data(), h2osoi_liq.size());
  write_raw(dir + "/out_state_zwt.f32", zwt.data(), zwt.size());
  write_raw(dir + "/out_state_plant_mass.f32", plant_mass.data(), plant_mass.size());
  printf("GPU launches %lld  H2D %.1f MB  D2H %.1f MB  step-kernel time %.1f ms\n",
         (long long)h9_launch_count(ctx), h9_h2d_bytes(ctx) / 1e6, h9_d2h_bytes(ctx) / 1e6, h9_step_kernel_ms(ctx));
  h9_destroy(ctx);
  printf(" CPU time (s)  %g\n", (double)(clock() - start) / CLOCKS_PER_SEC); /* HYBRID9.f90:572-573 */
  return 0;
}
