/*
 * h9_twin.cpp -- TEST-ONLY host build of the kernel source.
 *
 * Compiles hybrid9_b200/csrc/h9_physics.h (the exact text the CUDA kernels
 * instantiate) for the CPU with -ffp-contract=off, once with the libm policy and once with the
 * portable exact-mode policy, so
 * that the -m "not gpu" suite can diff the kernel's logic against the oracle
 * before any GPU time is spent.  Same libm as the oracle => the two must agree
 * bit for bit; a difference is a logic bug in one of the two restatements.
 * Never linked into libh9gpu.so and never used by the product path.
 */
#include <cstdint>
#include <cstring>

#include "../../hybrid9_b200/csrc/h9_physics.h"

using namespace h9;

/* All per-layer arrays are compact [ncell][8]; forcing is [ndays][7][ncell].
 * State arrays are updated in place. theta_out [ncell][8] gets the end state's
 * diagnostic theta.  daily_* (optional, [ndays][ncell]) get per-day npp, w_i, fT. */
template <class M>
static int twin_run(int nsteps, int ncell, int ndays, int nisurf, const float* zi, float* h2o, float* smp,
            float* rootr, const float* theta_s, const float* hksat, const float* bsw,
            const float* psi_s, const float* fmax, float* zwt, float* wa, float* lai,
            float* lai_litter, float* plant_mass, float* plant_foliage_mass, float* plant_length,
            float* rdepth, const int32_t* nplants, float* rnf_sum, const float* forcing,
            int do_grow, uint32_t* fault, float* theta_out, float* last_tran, float* last_evap,
            float* last_imb, int32_t* last_jwt, float* daily_npp, float* daily_wi,
            float* daily_ft) {
  Geo g;
  geo_init(g, zi, nisurf);
  for (int c = 0; c < ncell; ++c) {
    Params p;
    State s;
    for (int i = 0; i < NL; ++i) {
      p.theta_s[i] = theta_s[c * NL + i];
      p.hksat[i] = hksat[c * NL + i];
      p.bsw[i] = bsw[c * NL + i];
      p.psi_s[i] = psi_s[c * NL + i];
      s.h2o[i] = h2o[c * NL + i];
      s.smp[i] = smp[c * NL + i];
      s.rootr[i] = rootr[c * NL + i];
    }
    p.fmax = fmax[c];
    s.zwt = zwt[c];
    s.wa = wa[c];
    s.lai = lai[c];
    s.lai_litter = lai_litter[c];
    s.plant_mass = plant_mass[c];
    s.plant_foliage_mass = plant_foliage_mass[c];
    s.plant_length = plant_length[c];
    s.rdepth = rdepth[c];
    s.rnf_sum = rnf_sum[c];
    s.nplants = nplants[c];
    uint32_t ft = 0;
    StepOut so;
    memset(&so, 0, sizeof(so));
    for (int d = 0; d < ndays; ++d) {
      const float* f0 = forcing + ((size_t)d * 7) * ncell + c;
      Forcing f = {f0[0], f0[(size_t)ncell], f0[(size_t)2 * ncell], f0[(size_t)3 * ncell],
                   f0[(size_t)4 * ncell], f0[(size_t)5 * ncell], f0[(size_t)6 * ncell]};
      Day day;
      day_setup<M>(g, f, s.lai, s.lai_litter, day);
      for (int ns = 0; ns < nsteps; ++ns) ft |= hydrology_step<M>(g, p, day, s, so);
      if (do_grow) {
        GrowOut go;
        grow_day<M>(g, day.tas, s, go);
        if (daily_npp) daily_npp[(size_t)d * ncell + c] = go.npp;
        if (daily_wi) daily_wi[(size_t)d * ncell + c] = go.w_i;
        if (daily_ft) daily_ft[(size_t)d * ncell + c] = go.fT;
      }
    }
    for (int i = 0; i < NL; ++i) {
      h2o[c * NL + i] = s.h2o[i];
      smp[c * NL + i] = s.smp[i];
      rootr[c * NL + i] = s.rootr[i];
      if (theta_out) theta_out[c * NL + i] = theta_diag<M>(g, s.h2o[i], i);
    }
    zwt[c] = s.zwt;
    wa[c] = s.wa;
    lai[c] = s.lai;
    lai_litter[c] = s.lai_litter;
    plant_mass[c] = s.plant_mass;
    plant_foliage_mass[c] = s.plant_foliage_mass;
    plant_length[c] = s.plant_length;
    rdepth[c] = s.rdepth;
    rnf_sum[c] = s.rnf_sum;
    if (fault) fault[c] = ft;
    if (last_tran) last_tran[c] = so.qflx_tran_veg_col;
    if (last_evap) last_evap[c] = so.qflx_evap_grnd;
    if (last_imb) last_imb[c] = so.imbalance;
    if (last_jwt) last_jwt[c] = so.jwt;
  }
  return 0;
}

extern "C" {

/* math: 0 = MathLibm (the C library's powf/expf/logf, as the oracle), 1 = MathExact
 * (the portable kernels the GPU's exact mode runs) */
int h9t_run(int math, int nsteps, int ncell, int ndays, int nisurf, const float* zi, float* h2o, float* smp,
            float* rootr, const float* theta_s, const float* hksat, const float* bsw,
            const float* psi_s, const float* fmax, float* zwt, float* wa, float* lai,
            float* lai_litter, float* plant_mass, float* plant_foliage_mass, float* plant_length,
            float* rdepth, const int32_t* nplants, float* rnf_sum, const float* forcing,
            int do_grow, uint32_t* fault, float* theta_out, float* last_tran, float* last_evap,
            float* last_imb, int32_t* last_jwt, float* daily_npp, float* daily_wi,
            float* daily_ft) {
  if (nsteps < 0) nsteps = nisurf; /* HYDROLOGY calls per day; dt stays 86400/nisurf */
  if (math == 0)
    return twin_run<MathLibm>(nsteps, ncell, ndays, nisurf, zi, h2o, smp, rootr, theta_s, hksat, bsw, psi_s,
                              fmax, zwt, wa, lai, lai_litter, plant_mass, plant_foliage_mass,
                              plant_length, rdepth, nplants, rnf_sum, forcing, do_grow, fault,
                              theta_out, last_tran, last_evap, last_imb, last_jwt, daily_npp,
                              daily_wi, daily_ft);
  return twin_run<MathExact>(nsteps, ncell, ndays, nisurf, zi, h2o, smp, rootr, theta_s, hksat, bsw, psi_s,
                             fmax, zwt, wa, lai, lai_litter, plant_mass, plant_foliage_mass,
                             plant_length, rdepth, nplants, rnf_sum, forcing, do_grow, fault,
                             theta_out, last_tran, last_evap, last_imb, last_jwt, daily_npp,
                             daily_wi, daily_ft);
}

/* spot checks of the portable kernels against libm */
float h9t_pow(float a, float b) { return MathExact::pow(a, b); }
float h9t_exp(float a) { return MathExact::exp(a); }
float h9t_log(float a) { return MathExact::log(a); }

} /* extern "C" */
