/*
 * h9_oracle.cpp -- CPU restatement of HYBRID9's HYDROLOGY + GROW + driver loop.
 *
 * TEST INFRASTRUCTURE ONLY (see h9_oracle.h).  PARITY UNPINNED by the
 * reference's own tests (it ships none); this file follows
 * the Fortran source text statement by statement, in the same operation order,
 * in the same precision (default REAL == float), with libm powf/expf/logf.
 *
 * All citations are file:line under /root/reference/SOURCE.  Fortran operator
 * precedence is made explicit with parentheses: a*b/c == (a*b)/c, a/b/c ==
 * (a/b)/c, and sums run left to right.  x**2 and x**4 (integer powers) are
 * products, as every Fortran compiler expands them.
 *
 * Build: oracle/Makefile (strict: -O2 -ffp-contract=off; fast: -O3 for the CPU
 * baseline; f64: -DH9O_DOUBLE for the rounding-noise floor).
 */
#include "h9_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

typedef h9o_real real;

/* A Fortran default-REAL literal, converted to the working precision.  The
 * double build keeps the float-rounded constants so that it measures the
 * rounding noise of the arithmetic, not of the constants. */
#define K(x) ((real)(x##f))

namespace {

/* ---- constants: SHARED.f90:294-367,506; CONTROL.f90:21 ------------------ */
const int nlayers = 8;            /* nsoil_layers_max, SHARED.f90:294; HYBRID9.f90:126 */
const real zero = K(0.0), one = K(1.0);                      /* SHARED.f90:308-309 */
const real rhow = K(1000.0);                                 /* SHARED.f90:319 */
const float mair_f = 28.9655f, mwat_f = 18.015f;             /* SHARED.f90:323,327 */
const float gasc_f = 8.314510f;                              /* SHARED.f90:331 */
const float rgas_f = 1000.0f * gasc_f / mair_f;              /* SHARED.f90:335 */
const float mrat_f = mwat_f / mair_f;                        /* SHARED.f90:343 */
const float bymrat_f = 1.0f / mrat_f;                        /* SHARED.f90:347 */
const float deltx_f = bymrat_f - 1.0f;                       /* SHARED.f90:351 */
const real gasc = (real)gasc_f, rgas = (real)rgas_f, deltx = (real)deltx_f;
const real stbo = K(5.67E-8);                                /* SHARED.f90:339 */
const real tf = K(273.16);                                   /* SHARED.f90:363 */
const real smpmin = K(-1.0E8);                               /* SHARED.f90:367 */
const real trunc_ = K(1.0E-8);                               /* SHARED.f90:506 */
const real cp = K(1010.0);                                   /* HYDROLOGY.f90:35 */
const real watmin = K(0.01);                                 /* HYDROLOGY.f90:135 */
const real sla1 = K(23.0E-3);                                /* INIT.f90:154 */
const real plot_area = K(1.0);                               /* SHARED.f90:75 */

enum { F_PIVOT1 = 1u, F_PIVOT2 = 2u, F_RSUB = 4u, F_IMBAL = 8u };

inline real rmin(real a, real b) { return a < b ? a : b; } /* Fortran MIN */
inline real rmax(real a, real b) { return a > b ? a : b; } /* Fortran MAX */
inline real rpow(real a, real b) { return std::pow(a, b); }
inline real rexp(real a) { return std::exp(a); }
inline real rlog(real a) { return std::log(a); }
inline real rabs(real a) { return std::fabs(a); }

struct Forcing { /* one day, one cell: the 7 PGF fields, READ_PGF.f90:24-109 */
  real tas, rlds, rsds, huss, ps, pr, rhs;
};

struct Weather { /* HYBRID9.f90:168-184 */
  real tak, rh, Rnet, PAR, forc_rain, lamb;
};

inline Weather derive_weather(const Forcing& f) {
  Weather w;
  w.tak = f.tas;                                                        /* :168 */
  w.rh = f.rhs;                                                         /* :169 */
  real t2 = f.tas * f.tas;
  real t4 = t2 * t2;                                                    /* tas ** 4 */
  w.Rnet = (K(0.92) * f.rsds + f.rlds) - stbo * t4;                     /* :170-171 */
  w.PAR = (K(0.92) * f.rsds) * K(2.3);                                  /* :173 */
  w.forc_rain = (K(1.0E3) * f.pr) / rhow;                               /* :178 */
  w.lamb = (K(2503.0) - K(2.386) * (w.tak - tf)) * K(1.0E3);            /* :184 */
  return w;
}

/* pointers to the data of one cell, each 0-based (layer I at [I-1]) */
struct Cell {
  real* h2osoi_liq; /* (8) SHARED.f90:459 */
  real* zwt;        /* SHARED.f90:466 */
  real* wa;         /* SHARED.f90:472 */
  real* LAI;
  real* LAI_litter;
  real* plant_mass;
  real* plant_foliage_mass;
  real* plant_length;
  real* rdepth;
  real* rootr_col; /* (9) */
  int32_t* nplants;
  real* smp; /* (8): per cell, or the shared scratch vector when smp_leak */
  const real* theta_s;
  const real* hksat;
  const real* bsw;
  const real* psi_s;
  const real* Fmax;
  real* rnf_sum; /* SHARED.f90:134 (module scalar; per cell here) */
};

struct Geometry {
  real zi[10]; /* zi(0:9), INIT.f90:105,202-204 */
  real dz[10]; /* dz(1:9) at [1..9], INIT.f90:252-254 */
  real zc[10]; /* zc(1:9) at [1..9], INIT.f90:255-257 */
  real dt;     /* INIT.f90:214 */
};

int find_jwt(real zwt, const real* zi) { /* HYDROLOGY.f90:499-508 (and :923-931,1000-1007,1110-1116) */
  int jwt = nlayers;
  for (int I = 1; I <= nlayers; ++I) {
    if (zwt <= (zi[I] / K(1000.0))) {
      jwt = I - 1;
      break;
    }
  }
  return jwt;
}

/* ---- HYDROLOGY.f90:141-1283, one cell, one sub-step --------------------- */
void hydrology(const Geometry& g, const Cell& c, const Forcing& f, const Weather& wx,
               h9o_step_diag* out) {
  const real dt = g.dt;
  const real* zi = g.zi;
  real dz[10], zc[10]; /* module arrays whose element 9 is rewritten each call (:645-650) */
  for (int I = 1; I <= 9; ++I) {
    dz[I] = g.dz[I];
    zc[I] = g.zc[I];
  }
#define H2O(I) c.h2osoi_liq[(I)-1]
#define THS(I) c.theta_s[(I)-1]
#define HKS(I) c.hksat[(I)-1]
#define BSW(I) c.bsw[(I)-1]
#define PSI(I) c.psi_s[(I)-1]
#define ROOTR(I) c.rootr_col[(I)-1]
#define SMP(I) c.smp[(I)-1]
  real& zwt = *c.zwt;
  real& wa = *c.wa;
  const real LAI = *c.LAI;
  const real LAI_litter = *c.LAI_litter;
  const real tak = wx.tak, rh = wx.rh, Rnet = wx.Rnet, PAR = wx.PAR, forc_rain = wx.forc_rain,
             lamb = wx.lamb;
  uint32_t fault = 0;

  real theta[10], eff_porosity[10], vol_eq[11], zq[11], hk[10], dhkdw[10], dsmpdw[10];
  real qin[11], qout[11], dqidw0[11], dqidw1[11], dqodw1[11], dqodw2[11];
  real amx[11], bmx[11], cmx[11], rmx[11], dwat2[11], GAM[11], rnff[11];

  /* :141-151 */
  real w0 = forc_rain * dt + wa;
  for (int I = 1; I <= nlayers; ++I) {
    w0 = w0 + H2O(I);
    theta[I] = H2O(I) / (dz[I] * rhow / K(1.0E3));
  }

  /* SLakeHydrology :161-167 */
  real qflx_prec_grnd_rain = forc_rain;
  real qflx_top_soil = qflx_prec_grnd_rain;

  /* SurfaceRunoff :182-212 */
  real hkdepth = one / K(2.5);
  real fff = K(1.0) / hkdepth;
  real wtfact = *c.Fmax;
  real fsat = wtfact * rexp(K(-0.5) * fff * zwt);
  real fcov = fsat;
  real qflx_surf = fcov * qflx_top_soil;
  real frac_h2osfc = zero; /* :224 */

  /* met :232-263 */
  real tsv = tak * (one + f.huss * deltx);
  real rho = f.ps / (rgas * tsv);
  real desatdT = (K(4098.0) * (K(0.6108) * rexp((K(17.27) * (tak - tf)) / (tak - tf + K(237.3))))) /
                 ((tak - tf + K(237.3)) * (tak - tf + K(237.3)));
  desatdT = desatdT * K(18.0) / (gasc * tak);
  real esat = K(0.6108) * rexp(K(17.27) * (tak - tf) / (tak - tf + K(237.3)));
  esat = esat * K(18.0) / (gasc * tak);
  real VDD = esat * (one - rh / K(100.0));
  real gamma = (cp * f.ps / (lamb * K(0.622))) * (K(18.0E-3) / (gasc * tak));

  /* beta from the PREVIOUS call's smp :269-276 */
  real beta_save = zero, beta;
  for (int I = 1; I <= nlayers; ++I) {
    beta = one - (SMP(I) - zc[I]) / (K(-150000.0));
    beta = rmin(one, beta);
    beta = rmax(zero, beta);
    beta_save = beta_save + ROOTR(I) * beta;
  }
  beta = beta_save;

  /* rsc :283-295 */
  real rsc;
  if ((LAI > zero) && (beta > zero) && (PAR > zero)) {
    rsc = (K(1.0) / (PAR / (PAR + K(300.0)))) * K(400.0) /
          (K(2.0) * LAI * beta * (rpow(K(2.8), -(K(80.0) * rmax(zero, VDD) / rho))));
  } else {
    rsc = K(1.0E6);
  }
  rsc = rmax(rsc, K(1.0) / ((LAI / K(2.7)) * K(0.9) / (rho * K(1.0E3) / K(18.0))));

  /* rac, raa, ras :302-318 */
  real rac, raa, ras;
  if (LAI > zero) {
    rac = K(25.0) / (K(2.0) * LAI);
  } else {
    rac = K(1.0E6);
  }
  if (LAI <= K(4.0)) {
    raa = K(0.25) * LAI * K(42.0) + K(0.25) * (K(4.0) - LAI) * K(34.0);
    ras = K(0.25) * LAI * K(128.0) + K(0.25) * (K(4.0) - LAI) * K(49.0);
  } else {
    raa = K(42.0);
    ras = K(128.0);
  }

  /* rss :325-331 */
  real rss;
  if (theta[1] <= K(0.15)) {
    rss = (K(10.0) + K(1000.0) * LAI_litter) * rexp(K(0.3563) * K(100.0) * (K(0.15) - theta[1]));
  } else {
    rss = (K(10.0) + K(1000.0) * LAI_litter * (K(1.0) - theta[1] / THS(1)));
  }

  /* energy balance :335-384 */
  real Rnets = Rnet * rexp(K(-0.7) * LAI);
  real G = K(0.2) * Rnets;
  real PMc = (desatdT * (Rnet - G) + (rho * cp * VDD - desatdT * rac * (Rnets - G)) / (raa + rac)) /
             (desatdT + gamma * (one + rsc / (raa + rac)));
  real PMs = (desatdT * (Rnet - G) + (rho * cp * VDD - desatdT * ras * (Rnet - Rnets)) / (raa + ras)) /
             (desatdT + gamma * (one + rss / (raa + ras)));
  real Ra = (desatdT + gamma) * raa;
  real Rs = (desatdT + gamma) * ras + gamma * rss;
  real Rc = (desatdT + gamma) * rac + gamma * rsc;
  real Cc = one / (one + Rc * Ra / (Rs * (Rc + Ra)));
  real Cs = one / (one + Rs * Ra / (Rc * (Rs + Ra)));
  real LE = Cc * PMc + Cs * PMs;
  real VDD0 = VDD + (desatdT * (Rnet - G) - (desatdT + gamma) * LE) * raa / (rho * cp);
  real LEc = (desatdT * (Rnet - Rnets) + rho * cp * VDD0 / rac) /
             (desatdT + gamma * (K(1.0) + rsc / rac));
  real LEs = (desatdT * (Rnets - G) + rho * cp * VDD0 / ras) /
             (desatdT + gamma * (K(1.0) + rss / ras));
  real qflx_tran_veg_col = LEc * K(1.0E3) / (rhow * lamb); /* :388 */
  real qflx_evap_grnd = LEs * K(1.0E3) / (rhow * lamb);    /* :389 */

  /* evaporation limit :396-400 (evap_max(2:8) of :407-409 is never used) */
  real evap_max1 = dz[1] * (theta[1] - watmin) / dt - qflx_tran_veg_col * ROOTR(1);
  evap_max1 = rmax(zero, evap_max1);
  qflx_evap_grnd = rmin(evap_max1, qflx_evap_grnd);
  real qflx_ev_h2osfc = zero; /* :418 */

  /* Infiltration :426-478 */
  for (int I = 1; I <= nlayers; ++I) eff_porosity[I] = rmax(K(0.01), THS(I));
  real qflx_evap = qflx_evap_grnd;
  real qflx_in_soil = (one - frac_h2osfc) * (qflx_top_soil - qflx_surf);
  real qflx_in_h2osfc = frac_h2osfc * (qflx_top_soil - qflx_surf);
  qflx_in_soil = qflx_in_soil - (one - frac_h2osfc) * qflx_evap;
  qflx_in_h2osfc = qflx_in_h2osfc - frac_h2osfc * qflx_ev_h2osfc;
  (void)qflx_in_h2osfc;
  real qinmax = (one - fsat) * rmin(rmin(HKS(1), HKS(2)), HKS(3)); /* MINVAL(hksat(1:3)) :458 */
  real qflx_infl_excess = rmax(zero, qflx_in_soil - (one - frac_h2osfc) * qinmax);
  real qflx_infl = qflx_in_soil - qflx_infl_excess;
  qflx_surf = qflx_surf + qflx_infl_excess;
  qflx_infl_excess = zero;

  /* SoilWater :492-508 */
  real zwtmm = K(1000.0) * zwt;
  int jwt = find_jwt(zwt, zi);
  const int jwt_soilwater = jwt;

  /* equilibrium profile :517-567 */
  real tempi, temp0, voleq1;
  for (int I = 1; I <= nlayers; ++I) {
    if (zwtmm <= zi[I - 1]) {
      vol_eq[I] = THS(I);
    } else if ((zwtmm < zi[I]) && (zwtmm > zi[I - 1])) {
      tempi = one;
      temp0 = rpow((((-PSI(I)) + zwtmm - zi[I - 1]) / (-PSI(I))), (one - one / BSW(I)));
      voleq1 = PSI(I) * THS(I) / (one - one / BSW(I)) / (zwtmm - zi[I - 1]) * (tempi - temp0);
      vol_eq[I] = (voleq1 * (zwtmm - zi[I - 1]) + THS(I) * (zi[I] - zwtmm)) / (zi[I] - zi[I - 1]);
      vol_eq[I] = rmin(THS(I), vol_eq[I]);
      vol_eq[I] = rmax(vol_eq[I], zero);
    } else {
      tempi = rpow(((-PSI(I) + zwtmm - zi[I]) / (-PSI(I))), (K(1.0) - K(1.0) / BSW(I)));
      temp0 = rpow(((-PSI(I) + zwtmm - zi[I - 1]) / (-PSI(I))), (K(1.0) - K(1.0) / BSW(I)));
      vol_eq[I] = PSI(I) * THS(I) / (K(1.0) - K(1.0) / BSW(I)) / (zi[I] - zi[I - 1]) * (tempi - temp0);
      vol_eq[I] = rmax(vol_eq[I], K(0.0));
      vol_eq[I] = rmin(THS(I), vol_eq[I]);
    }
    zq[I] = PSI(I) * rpow(rmax(vol_eq[I] / THS(I), K(0.01)), (-BSW(I)));
    zq[I] = rmax(smpmin, zq[I]);
  }

  /* 9th layer :574-590 */
  {
    const int I = nlayers;
    if (jwt == nlayers) {
      tempi = K(1.0);
      temp0 = rpow(((-PSI(I) + zwtmm - zi[I]) / (-PSI(I))), (K(1.0) - K(1.0) / BSW(I)));
      vol_eq[I + 1] = PSI(I) * THS(I) / (K(1.0) - K(1.0) / BSW(I)) / (zwtmm - zi[I]) * (tempi - temp0);
      vol_eq[I + 1] = rmax(vol_eq[I + 1], K(0.0));
      vol_eq[I + 1] = rmin(THS(I), vol_eq[I + 1]);
      zq[I + 1] = PSI(I) * rpow(rmax(vol_eq[I + 1] / THS(I), K(0.01)), (-BSW(I)));
      zq[I + 1] = rmax(smpmin, zq[I + 1]);
    } else {
      zq[I + 1] = zero; /* stale module scratch in the reference; never read on this path */
    }
  }

  /* hk, dhkdw, smp, dsmpdw :598-639 */
  real s1, s2, s_node;
  for (int I = 1; I <= nlayers; ++I) {
    const int Ip = std::min(nlayers, I + 1);
    s1 = K(0.5) * (theta[I] + theta[Ip]) / (K(0.5) * (THS(I) + THS(Ip)));
    s1 = rmin(one, s1);
    s2 = HKS(I) * rpow(s1, (K(2.0) * BSW(I) + K(2.0)));
    hk[I] = s1 * s2;
    dhkdw[I] = (K(2.0) * BSW(I) + K(3.0)) * s2 * (one / (THS(I) + THS(Ip)));
    s_node = rmax(theta[I] / THS(I), K(0.01));
    s_node = rmin(one, s_node);
    SMP(I) = PSI(I) * rpow(s_node, (-BSW(I)));
    SMP(I) = rmax(smpmin, SMP(I));
    dsmpdw[I] = (-BSW(I)) * SMP(I) / (s_node * THS(I));
  }

  /* aquifer node :645-650 */
  zc[nlayers + 1] = K(0.5) * (zwtmm + zc[nlayers]);
  if (jwt < nlayers) {
    dz[nlayers + 1] = dz[nlayers];
  } else {
    dz[nlayers + 1] = zwtmm - zc[nlayers];
  }

  /* tridiagonal rows :661-799 */
  real den, dzq, num, smp1, dsmpdw1;
  {
    const int I = 1;
    qin[I] = qflx_infl;
    den = (zc[I + 1] - zc[I]);
    dzq = (zq[I + 1] - zq[I]);
    num = (SMP(I + 1) - SMP(I)) - dzq;
    qout[I] = -hk[I] * num / den;
    dqodw1[I] = -(-hk[I] * dsmpdw[I] + num * dhkdw[I]) / den;
    dqodw2[I] = -(hk[I] * dsmpdw[I + 1] + num * dhkdw[I]) / den;
    rmx[I] = qin[I] - qout[I] - qflx_tran_veg_col * ROOTR(I);
    amx[I] = zero;
    bmx[I] = dz[I] / dt + dqodw1[I];
    cmx[I] = dqodw2[I];
  }
  for (int I = 2; I <= nlayers - 1; ++I) {
    den = zc[I] - zc[I - 1];
    dzq = zq[I] - zq[I - 1];
    num = SMP(I) - SMP(I - 1) - dzq;
    qin[I] = -hk[I - 1] * num / den;
    dqidw0[I] = -(-hk[I - 1] * dsmpdw[I - 1] + num * dhkdw[I - 1]) / den;
    dqidw1[I] = -(hk[I - 1] * dsmpdw[I] + num * dhkdw[I - 1]) / den;
    den = zc[I + 1] - zc[I];
    dzq = zq[I + 1] - zq[I];
    num = (SMP(I + 1) - SMP(I)) - dzq;
    qout[I] = -hk[I] * num / den;
    dqodw1[I] = -(-hk[I] * dsmpdw[I] + num * dhkdw[I]) / den;
    dqodw2[I] = -(hk[I] * dsmpdw[I + 1] + num * dhkdw[I]) / den;
    rmx[I] = qin[I] - qout[I] - qflx_tran_veg_col * ROOTR(I);
    amx[I] = -dqidw0[I];
    bmx[I] = dz[I] / dt - dqidw1[I] + dqodw1[I];
    cmx[I] = dqodw2[I];
  }
  {
    const int I = nlayers;
    if (I > jwt) { /* water table in the soil column :712-735 */
      den = zc[I] - zc[I - 1];
      dzq = zq[I] - zq[I - 1];
      num = SMP(I) - SMP(I - 1) - dzq;
      qin[I] = -hk[I - 1] * num / den;
      dqidw0[I] = -(-hk[I - 1] * dsmpdw[I - 1] + num * dhkdw[I - 1]) / den;
      dqidw1[I] = -(hk[I - 1] * dsmpdw[I] + num * dhkdw[I - 1]) / den;
      qout[I] = zero;
      dqodw1[I] = zero;
      rmx[I] = qin[I] - qout[I] - qflx_tran_veg_col * ROOTR(I);
      amx[I] = -dqidw0[I];
      bmx[I] = dz[I] / dt - dqidw1[I] + dqodw1[I];
      cmx[I] = zero;
      rmx[I + 1] = zero;
      amx[I + 1] = zero;
      bmx[I + 1] = dz[I + 1] / dt;
      cmx[I + 1] = zero;
    } else { /* water table below the soil column :737-799 */
      s_node = rmax(K(0.5) * (one + theta[I] / THS(I)), K(0.01));
      s_node = rmin(one, s_node);
      smp1 = PSI(I) * rpow(s_node, (-BSW(I)));
      smp1 = rmax(smpmin, smp1);
      dsmpdw1 = -BSW(I) * smp1 / (s_node * THS(I));
      den = zc[I] - zc[I - 1];
      dzq = zq[I] - zq[I - 1];
      num = SMP(I) - SMP(I - 1) - dzq;
      qin[I] = -hk[I - 1] * num / den;
      dqidw0[I] = -(-hk[I - 1] * dsmpdw[I - 1] + num * dhkdw[I - 1]) / den;
      dqidw1[I] = -(hk[I - 1] * dsmpdw[I] + num * dhkdw[I - 1]) / den;
      den = zc[I + 1] - zc[I];
      dzq = zq[I + 1] - zq[I];
      num = smp1 - SMP(I) - dzq;
      qout[I] = -hk[I] * num / den;
      dqodw1[I] = -(-hk[I] * dsmpdw[I] + num * dhkdw[I]) / den;
      dqodw2[I] = -(hk[I] * dsmpdw1 + num * dhkdw[I]) / den;
      rmx[I] = qin[I] - qout[I] - qflx_tran_veg_col * ROOTR(I);
      amx[I] = -dqidw0[I];
      bmx[I] = dz[I] / dt - dqidw1[I] + dqodw1[I];
      cmx[I] = dqodw2[I];
      qin[I + 1] = qout[I];
      dqidw0[I + 1] = -(-hk[I] * dsmpdw[I] + num * dhkdw[I]) / den;
      dqidw1[I + 1] = -(hk[I] * dsmpdw1 + num * dhkdw[I]) / den;
      qout[I + 1] = zero;
      dqodw1[I + 1] = zero;
      rmx[I + 1] = qin[I + 1] - qout[I + 1];
      amx[I + 1] = -dqidw0[I + 1];
      bmx[I + 1] = dz[I + 1] / dt - dqidw1[I + 1] + dqodw1[I + 1];
      cmx[I + 1] = zero;
    }
  }

  /* Thomas :806-837 (the reference STOPs on a zero pivot; we flag and go on) */
  if (bmx[1] == K(0.0)) fault |= F_PIVOT1;
  real BET = bmx[1];
  dwat2[1] = rmx[1] / BET;
  for (int I = 2; I <= nlayers + 1; ++I) {
    GAM[I] = cmx[I - 1] / BET;
    BET = bmx[I] - amx[I] * GAM[I];
    if (BET == K(0.0)) fault |= F_PIVOT2;
    dwat2[I] = (rmx[I] - amx[I] * dwat2[I - 1]) / BET;
  }
  for (int I = nlayers; I >= 1; --I) dwat2[I] = dwat2[I] - GAM[I + 1] * dwat2[I + 1];

  /* :845-850 */
  for (int I = 1; I <= nlayers; ++I) H2O(I) = H2O(I) + dwat2[I] * dz[I];

  /* recharge :856-904 */
  real qcharge;
  if (jwt < nlayers) {
    real wh_zwt = zero;
    s_node = rmax(theta[jwt + 1] / THS(jwt + 1), K(0.01));
    s1 = rmin(one, s_node);
    real ka = HKS(jwt + 1) * rpow(s1, (K(2.0) * BSW(jwt + 1) + K(3.0)));
    smp1 = rmax(smpmin, SMP(std::max(1, jwt)));
    real wh = smp1 - zq[std::max(1, jwt)];
    if (jwt == 0) {
      qcharge = -ka * (wh_zwt - wh) / (zwtmm + one);
    } else {
      qcharge = -ka * (wh_zwt - wh) / ((zwtmm - zc[jwt]) * K(2.0));
    }
    qcharge = rmax(K(-10.0) / dt, qcharge);
    qcharge = rmin(K(10.0) / dt, qcharge);
  } else {
    qcharge = dwat2[nlayers + 1] * dz[nlayers + 1] / dt;
  }

  /* Drainage :923-940 */
  jwt = find_jwt(zwt, zi);
  real rous = THS(nlayers) * (one - rpow((one + zwtmm / (-PSI(nlayers))), (-one / BSW(nlayers))));
  rous = rmax(rous, K(0.02));

  real s_y, qcharge_tot, qcharge_layer;
  if (jwt == nlayers) { /* :946-951 */
    wa = wa + qcharge * dt;
    zwt = zwt - (qcharge * dt) / K(1000.0) / rous;
  } else { /* :953-1009; zwtmm is deliberately the stale value of :492 */
    qcharge_tot = qcharge * dt;
    if (qcharge_tot > zero) {
      for (int I = jwt + 1; I >= 1; --I) {
        s_y = THS(I) * (one - rpow((one + zwtmm / (-PSI(I))), (-one / BSW(I))));
        s_y = rmax(s_y, K(0.02));
        qcharge_layer = rmin(qcharge_tot, s_y * (zwtmm - zi[I - 1]));
        qcharge_layer = rmax(qcharge_layer, zero);
        if (s_y > zero) zwt = zwt - qcharge_layer / s_y / K(1000.0);
        qcharge_tot = qcharge_tot - qcharge_layer;
        if (qcharge_tot <= zero) break;
      }
    } else {
      for (int I = jwt + 1; I <= nlayers; ++I) {
        s_y = THS(I) * (one - rpow((one + zwtmm / (-PSI(I))), (-one / BSW(I))));
        s_y = rmax(s_y, K(0.02));
        qcharge_layer = rmax(qcharge_tot, -s_y * (zi[I] - zwtmm));
        qcharge_layer = rmin(qcharge_layer, zero);
        qcharge_tot = qcharge_tot - qcharge_layer;
        if (qcharge_tot >= zero) {
          zwt = zwt - qcharge_layer / s_y / K(1000.0);
          break;
        } else {
          zwt = zi[I] / K(1000.0);
        }
      }
      if (qcharge_tot > zero) zwt = zwt - qcharge_tot / K(1000.0) / rous;
    }
    jwt = find_jwt(zwt, zi); /* :1000-1007 */
  }

  zwtmm = K(1000.0) * zwt; /* :1015 */

  /* baseflow :1024-1035 */
  real rsub_top_max = K(5.5E-3);
  real rsub_top = rsub_top_max * rexp(-fff * zwt);
  rous = THS(nlayers) * (one - rpow((one + zwtmm / (-PSI(nlayers))), (-one / BSW(nlayers))));
  rous = rmax(rous, K(0.02));
  for (int I = 1; I <= nlayers + 1; ++I) rnff[I] = K(0.0);

  if (jwt == nlayers) { /* :1048-1058; jwt is NOT recomputed on this path */
    wa = wa - rsub_top * dt;
    zwt = zwt + (rsub_top * dt) / K(1000.0) / rous;
    H2O(nlayers) = H2O(nlayers) + rmax(K(0.0), (wa - K(5000.0)));
    wa = rmin(wa, K(5000.0));
    rnff[nlayers + 1] = rsub_top;
  } else { /* :1060-1118 */
    real rsub_top_tot = -rsub_top * dt;
    if (rsub_top_tot > zero) {
      fault |= F_RSUB; /* reference STOPs :1068-1071 */
    } else {
      real rsub_top_layer;
      for (int I = jwt + 1; I <= nlayers; ++I) {
        s_y = THS(I) * (one - rpow((one + zwtmm / (-PSI(I))), (-one / BSW(I))));
        s_y = rmax(s_y, K(0.02));
        rsub_top_layer = rmax(rsub_top_tot, -(s_y * (zi[I] - zwtmm)));
        rsub_top_layer = rmin(rsub_top_layer, zero);
        H2O(I) = H2O(I) + rsub_top_layer;
        rnff[I] = -rsub_top_layer;
        rsub_top_tot = rsub_top_tot - rsub_top_layer;
        if (rsub_top_tot >= zero) {
          zwt = zwt - rsub_top_layer / s_y / K(1000.0);
          break;
        } else {
          zwt = zi[I] / K(1000.0);
        }
      }
      /* residual, unconditional :1100-1102 */
      zwt = zwt - rsub_top_tot / K(1000.0) / rous;
      wa = wa + rsub_top_tot;
      rnff[nlayers + 1] = rnff[nlayers + 1] - rsub_top_tot;
    }
    jwt = find_jwt(zwt, zi); /* :1110-1116 */
  }

  zwt = rmax(K(0.0), zwt);  /* :1122 */
  zwt = rmin(K(80.0), zwt); /* :1123 */

  /* excess cascade :1131-1152 */
  real xsi;
  for (int I = nlayers; I >= 2; --I) {
    xsi = rmax(H2O(I) - eff_porosity[I] * dz[I], zero);
    H2O(I) = rmin(eff_porosity[I] * dz[I], H2O(I));
    H2O(I - 1) = H2O(I - 1) + xsi;
  }
  real xs1 = rmax(rmax(H2O(1), zero) - rmax(zero, (THS(1) * dz[1])), zero);
  H2O(1) = rmin(rmax(zero, THS(1) * dz[1]), H2O(1));
  real qflx_rsub_sat = xs1 / dt;

  /* dryness repair :1161-1205 */
  real xs;
  for (int I = 1; I <= nlayers - 1; ++I) {
    if (H2O(I) < watmin) {
      xs = watmin - H2O(I);
      if (I == jwt) zwt = zwt + xs / eff_porosity[I] / K(1000.0);
    } else {
      xs = zero;
    }
    H2O(I) = H2O(I) + xs;
    H2O(I + 1) = H2O(I + 1) - xs;
  }
  {
    const int I = nlayers;
    if (H2O(I) < watmin) {
      xs = watmin - H2O(I);
      for (int J = nlayers - 1; J >= 1; --J) {
        real available_h2osoi_liq = rmax(H2O(J) - watmin - xs, zero);
        if (available_h2osoi_liq >= xs) {
          H2O(I) = H2O(I) + xs;
          H2O(J) = H2O(J) - xs;
          xs = zero;
          break;
        } else {
          H2O(I) = H2O(I) + available_h2osoi_liq;
          H2O(J) = H2O(J) - available_h2osoi_liq;
          xs = xs - available_h2osoi_liq;
        }
      }
    } else {
      xs = zero;
    }
    H2O(I) = H2O(I) + xs; /* :1205 */
  }
  rsub_top = rsub_top - xs / dt; /* :1211 */

  /* balance :1221-1236 */
  real w1 = ((K(1.0) - frac_h2osfc) * (qflx_surf + qflx_evap_grnd + qflx_tran_veg_col) + rsub_top +
             qflx_rsub_sat) * dt + wa;
  for (int I = 1; I <= nlayers; ++I) {
    w1 = w1 + H2O(I);
    theta[I] = rmax(H2O(I), K(1.0E-6)) / (dz[I] * rhow / K(1000.0));
  }
  if (rabs(w1 - w0) > K(0.1)) fault |= F_IMBAL; /* :1244 */
  /* written so that a NaN imbalance is also flagged */
  if (!(rabs(w1 - w0) <= K(0.1))) fault |= F_IMBAL;

  /* :1282-1283 */
  *c.rnf_sum = *c.rnf_sum + qflx_surf * dt;
  *c.rnf_sum = *c.rnf_sum + rsub_top * dt;

  if (out) {
    for (int I = 1; I <= nlayers; ++I) out->theta[I - 1] = theta[I];
    out->qflx_tran_veg_col = qflx_tran_veg_col;
    out->qflx_evap_grnd = qflx_evap_grnd;
    out->qflx_surf = qflx_surf;
    out->rsub_top = rsub_top;
    out->qflx_rsub_sat = qflx_rsub_sat;
    out->qflx_infl = qflx_infl;
    out->qcharge = qcharge;
    out->fsat = fsat;
    out->beta = beta;
    out->rsc = rsc;
    out->w0 = w0;
    out->w1 = w1;
    out->rnf_inc = qflx_surf * dt + rsub_top * dt;
    out->jwt_soilwater = jwt_soilwater;
    out->jwt_final = jwt;
    out->fault = fault;
  }
#undef H2O
#undef THS
#undef HKS
#undef BSW
#undef PSI
#undef ROOTR
#undef SMP
}

/* root profile, shared by INIT.f90:791-797 and GROW.f90:176-182 */
void add_root_profile(const real* zi, real rdepth, real* rootr_col) {
  real decay = rexp(rlog(K(0.1)) / (rdepth / K(10.0)));
  for (int I = 1; I <= nlayers; ++I) {
    rootr_col[I - 1] = rootr_col[I - 1] + (K(1.0) - rpow(decay, (zi[I] / K(10.0)))) -
                       (K(1.0) - rpow(decay, (zi[I - 1] / K(10.0))));
  }
}

struct GrowDiag {
  real npp, w_i, fT;
};

/* ---- GROW.f90:55-201, one cell, one day --------------------------------- */
void grow(const Geometry& g, const Cell& c, real tas, GrowDiag* out) {
  real w_i_save = zero, w_i, fT;
  for (int I = 1; I <= nlayers; ++I) { /* :55-62 */
    w_i = (K(-150000.0) - c.smp[I - 1]) / (K(-150000.0) - (K(-50000.0)));
    w_i = rmax(zero, w_i);
    w_i = rmin(one, w_i);
    w_i_save = w_i_save + c.rootr_col[I - 1] * w_i;
  }
  w_i = w_i_save;
  if ((tas - tf) > K(18.0)) { /* :66-72 */
    real q = rabs(tas - tf - K(18.0)) / K(21.0);
    fT = one - q * q;
  } else {
    real q = rabs(tas - tf - K(18.0)) / K(25.0);
    fT = one - q * q;
    fT = rmax(zero, fT);
    fT = rmin(one, fT);
  }
  for (int I = 0; I < 9; ++I) c.rootr_col[I] = zero; /* :76 */
  real npp = zero;                                   /* :78 */
  const int np = *c.nplants;
  for (int Kp = 1; Kp <= np; ++Kp) { /* nplants_max == 1, SHARED.f90:63 */
    real grow_plant_mass = (K(1000.0) / K(365.0)) * w_i * fT; /* :90 */
    real grow_foliage_mass = grow_plant_mass / K(3.3);        /* :91 */
    real loss_plant_mass = (K(0.1) / K(365.0)) * c.plant_mass[Kp - 1]; /* :134 */
    real loss_foliage_mass =
        (K(1.0) / K(365.0)) * c.plant_foliage_mass[Kp - 1] / rmin(one, rmax(K(0.01), w_i)); /* :136 */
    if (w_i < K(0.6)) loss_foliage_mass = K(0.1) * c.plant_foliage_mass[Kp - 1];            /* :138 */
    real dplant_mass = grow_plant_mass - loss_plant_mass;
    real dplant_foliage_mass = grow_foliage_mass - loss_foliage_mass;
    c.plant_mass[Kp - 1] = c.plant_mass[Kp - 1] + dplant_mass;
    c.plant_foliage_mass[Kp - 1] = c.plant_foliage_mass[Kp - 1] + dplant_foliage_mass;
    c.plant_length[Kp - 1] =
        rpow((K(400.0) * c.plant_mass[Kp - 1] / K(3.142E-3)), (one / K(3.0))); /* :155 */
    real dLAI = dplant_foliage_mass * sla1;                                     /* :161 */
    *c.LAI = *c.LAI + dLAI;
    *c.LAI = rmax(K(0.001), *c.LAI);
    *c.LAI_litter = *c.LAI_litter + rmax(zero, dLAI); /* :167 */
    c.rdepth[Kp - 1] = K(0.3) * c.plant_length[Kp - 1]; /* :171 */
    add_root_profile(g.zi, c.rdepth[Kp - 1], c.rootr_col); /* :176-182 */
    npp = npp + dplant_mass;                               /* :186 */
  }
  if (np > 1) { /* :194 (unreachable with nplants_max == 1) */
    for (int I = 0; I < 9; ++I) c.rootr_col[I] = c.rootr_col[I] / (real)np;
  }
  *c.LAI_litter = *c.LAI_litter - K(0.02) * *c.LAI_litter; /* :201 */
  if (out) {
    out->npp = npp;
    out->w_i = w_i;
    out->fT = fT;
  }
}

} /* namespace */

struct h9o_ctx {
  int lon_c = 0, lat_c = 0, nisurf = 48, nyr = 1;
  Geometry geo;
  bool configured = false, have_soil = false;
  int loop_order = 0, smp_leak = 0, nthreads = 1, real_evap = 0;
  std::vector<real> evap_sum;
  size_t ncell = 0;
  std::vector<int32_t> soil_tex, nplants, land;
  std::vector<real> theta_s, hksat, bsw, psi_s, fmax;
  std::vector<real> h2osoi_liq, zwt, wa, lai, lai_litter, plant_mass, plant_foliage_mass,
      plant_length, rdepth, rootr_col, smp;
  real smp_shared[8]; /* SHARED.f90:198: one scratch vector for all cells */
  /* annual sums, HYBRID9.f90:134-146 (module scalars there; per cell here so that
   * both loop orders share the code) */
  std::vector<real> npp_sum, plant_mass_sum, rnf_sum, h2osoi_sum_total, theta_sum;
  std::vector<int32_t> nt, cur_year;
  /* axy_*: (lon_c,lat_c,NYR) and (8,lon_c,lat_c,NYR), INIT.f90:325-340,402-414 */
  std::vector<real> axy_npp, axy_plant_mass, axy_rnf, axy_evap, axy_theta_total, axy_theta;
  /* faults: sticky word per cell and the first fault of each cell */
  std::vector<uint32_t> fault, first_code;
  std::vector<int32_t> first_day, first_sub;
  std::vector<real> first_imb;
  int64_t day_counter = 0;
  std::vector<h9o_step_diag> last_diag; /* per land cell, filled by h9o_hydrology_step */

  Cell cell(size_t ci) {
    Cell c;
    c.h2osoi_liq = &h2osoi_liq[8 * ci];
    c.zwt = &zwt[ci];
    c.wa = &wa[ci];
    c.LAI = &lai[ci];
    c.LAI_litter = &lai_litter[ci];
    c.plant_mass = &plant_mass[ci];
    c.plant_foliage_mass = &plant_foliage_mass[ci];
    c.plant_length = &plant_length[ci];
    c.rdepth = &rdepth[ci];
    c.rootr_col = &rootr_col[9 * ci];
    c.nplants = &nplants[ci];
    c.smp = smp_leak ? smp_shared : &smp[8 * ci];
    c.theta_s = &theta_s[8 * ci];
    c.hksat = &hksat[8 * ci];
    c.bsw = &bsw[8 * ci];
    c.psi_s = &psi_s[8 * ci];
    c.Fmax = &fmax[ci];
    c.rnf_sum = &rnf_sum[ci];
    return c;
  }
};

namespace {

void note_fault(h9o_ctx* ctx, size_t ci, uint32_t code, int day, int sub, real imb) {
  if (!code) return;
  if (!ctx->fault[ci]) {
    ctx->first_code[ci] = code;
    ctx->first_day[ci] = day;
    ctx->first_sub[ci] = sub;
    ctx->first_imb[ci] = imb;
  }
  ctx->fault[ci] |= code;
}

/* one day of one cell: HYBRID9.f90:150-254 plus the year-end means :263-291 */
void cell_day(h9o_ctx* ctx, size_t ci, const Forcing& f, int iy, int day_abs) {
  Cell c = ctx->cell(ci);
  if (ctx->cur_year[ci] != iy) { /* start of a year: :134-146 */
    ctx->cur_year[ci] = iy;
    ctx->nt[ci] = 0;
    ctx->npp_sum[ci] = zero;
    ctx->plant_mass_sum[ci] = zero;
    ctx->rnf_sum[ci] = zero;
    ctx->evap_sum[ci] = zero;
    ctx->h2osoi_sum_total[ci] = zero;
    for (int I = 0; I < 8; ++I) ctx->theta_sum[8 * ci + I] = zero;
  }
  Weather wx = derive_weather(f);
  h9o_step_diag d;
  for (int NS = 1; NS <= ctx->nisurf; ++NS) { /* :193-211 */
    hydrology(ctx->geo, c, f, wx, &d);
    note_fault(ctx, ci, d.fault, day_abs, NS, d.w1 - d.w0);
    if (ctx->real_evap) ctx->evap_sum[ci] = ctx->evap_sum[ci] + (d.qflx_evap_grnd + d.qflx_tran_veg_col);
  }
  GrowDiag gd;
  grow(ctx->geo, c, f.tas, &gd); /* :217 */
  /* :242-253 */
  for (int Kp = 1; Kp <= *c.nplants; ++Kp)
    ctx->plant_mass_sum[ci] = ctx->plant_mass_sum[ci] + c.plant_mass[Kp - 1];
  ctx->npp_sum[ci] = ctx->npp_sum[ci] + gd.npp;
  for (int I = 0; I < 8; ++I) {
    ctx->theta_sum[8 * ci + I] = ctx->theta_sum[8 * ci + I] + d.theta[I];
    ctx->h2osoi_sum_total[ci] = ctx->h2osoi_sum_total[ci] + c.h2osoi_liq[I];
  }
  ctx->nt[ci] += 1;
  /* :263-291, refreshed after every day so that partial years are defined;
   * after the last day of a year these are the reference's values */
  if (iy >= 1 && iy <= ctx->nyr) {
    const size_t o = (size_t)(iy - 1) * ctx->ncell + ci;
    const int nt = ctx->nt[ci];
    ctx->axy_npp[o] = ctx->npp_sum[ci];
    ctx->axy_plant_mass[o] = ctx->plant_mass_sum[ci] / (real)nt;
    ctx->axy_rnf[o] = ctx->rnf_sum[ci] / (real)(nt * ctx->nisurf);
    /* evap_sum is never accumulated in the reference (:137,276): 0 unless real_evap is on */
    ctx->axy_evap[o] = (ctx->real_evap ? ctx->evap_sum[ci] : zero) / (real)(nt * ctx->nisurf);
    for (int I = 0; I < 8; ++I)
      ctx->axy_theta[8 * o + I] = ctx->theta_sum[8 * ci + I] / (real)nt;
    ctx->axy_theta_total[o] = ctx->h2osoi_sum_total[ci] / (real)nt;
  }
}

Forcing forcing_at(const h9o_ctx* ctx, size_t ci, int d, const real* tas, const real* rlds,
                   const real* rsds, const real* huss, const real* ps, const real* pr,
                   const real* rhs) {
  const size_t o = (size_t)d * ctx->ncell + ci; /* (x,y,iT), x fastest */
  Forcing f = {tas[o], rlds[o], rsds[o], huss[o], ps[o], pr[o], rhs[o]};
  return f;
}

} /* namespace */

extern "C" {

int h9o_sizeof_real(void) { return (int)sizeof(real); }

int h9o_create(h9o_ctx** ctx) {
  if (!ctx) return -1;
  *ctx = new h9o_ctx();
  return 0;
}

int h9o_destroy(h9o_ctx* ctx) {
  delete ctx;
  return 0;
}

int h9o_configure(h9o_ctx* ctx, int lon_c, int lat_c, int nisurf, const h9o_real zi[10], int nyr) {
  if (!ctx || lon_c < 1 || lat_c < 1 || nisurf < 1 || !zi || nyr < 1) return -1;
  ctx->lon_c = lon_c;
  ctx->lat_c = lat_c;
  ctx->nisurf = nisurf;
  ctx->nyr = nyr;
  ctx->ncell = (size_t)lon_c * lat_c;
  Geometry& g = ctx->geo;
  for (int I = 0; I <= 9; ++I) g.zi[I] = zi[I];
  g.dt = K(86400.0) / (real)nisurf; /* INIT.f90:214 */
  g.dz[0] = g.zc[0] = zero;
  for (int I = 1; I <= 9; ++I) g.dz[I] = g.zi[I] - g.zi[I - 1];      /* INIT.f90:252-254 */
  for (int I = 1; I <= 9; ++I) g.zc[I] = g.zi[I] - g.dz[I] / K(2.0); /* INIT.f90:255-257 */
  const size_t n = ctx->ncell;
  ctx->soil_tex.assign(n, 0);
  ctx->nplants.assign(n, 0);
  ctx->theta_s.assign(8 * n, zero);
  ctx->hksat.assign(8 * n, zero);
  ctx->bsw.assign(8 * n, zero);
  ctx->psi_s.assign(8 * n, zero);
  ctx->fmax.assign(n, zero);
  ctx->h2osoi_liq.assign(8 * n, zero);
  ctx->zwt.assign(n, zero);
  ctx->wa.assign(n, zero);
  ctx->lai.assign(n, zero);
  ctx->lai_litter.assign(n, zero);
  ctx->plant_mass.assign(n, zero);
  ctx->plant_foliage_mass.assign(n, zero);
  ctx->plant_length.assign(n, zero);
  ctx->rdepth.assign(n, zero);
  ctx->rootr_col.assign(9 * n, zero);
  ctx->smp.assign(8 * n, zero);
  for (int I = 0; I < 8; ++I) ctx->smp_shared[I] = zero;
  ctx->npp_sum.assign(n, zero);
  ctx->plant_mass_sum.assign(n, zero);
  ctx->rnf_sum.assign(n, zero);
  ctx->evap_sum.assign(n, zero);
  ctx->h2osoi_sum_total.assign(n, zero);
  ctx->theta_sum.assign(8 * n, zero);
  ctx->nt.assign(n, 0);
  ctx->cur_year.assign(n, 0);
  const real nan = std::nan("");
  const size_t ny = (size_t)nyr * n; /* fills: INIT.f90:402-414 */
  ctx->axy_npp.assign(ny, nan);
  ctx->axy_plant_mass.assign(ny, nan);
  ctx->axy_rnf.assign(ny, nan);
  ctx->axy_evap.assign(ny, nan);
  ctx->axy_theta.assign(8 * ny, nan);
  ctx->axy_theta_total.assign(ny, zero);
  ctx->fault.assign(n, 0);
  ctx->first_code.assign(n, 0);
  ctx->first_day.assign(n, 0);
  ctx->first_sub.assign(n, 0);
  ctx->first_imb.assign(n, zero);
  ctx->day_counter = 0;
  ctx->land.clear();
  ctx->configured = true;
  ctx->have_soil = false;
  return 0;
}

int h9o_set_soil(h9o_ctx* ctx, const int32_t* soil_tex, const h9o_real* theta_s,
                 const h9o_real* hksat, const h9o_real* bsw, const h9o_real* psi_s,
                 const h9o_real* fmax) {
  if (!ctx || !ctx->configured || !soil_tex || !theta_s || !hksat || !bsw || !psi_s || !fmax)
    return -1;
  const size_t n = ctx->ncell;
  std::copy(soil_tex, soil_tex + n, ctx->soil_tex.begin());
  std::copy(theta_s, theta_s + 8 * n, ctx->theta_s.begin());
  std::copy(hksat, hksat + 8 * n, ctx->hksat.begin());
  std::copy(bsw, bsw + 8 * n, ctx->bsw.begin());
  std::copy(psi_s, psi_s + 8 * n, ctx->psi_s.begin());
  std::copy(fmax, fmax + n, ctx->fmax.begin());
  ctx->land.clear();
  /* HYBRID9.f90:120-123: y outer, x inner; SUM in index order */
  for (int y = 1; y <= ctx->lat_c; ++y) {
    for (int x = 1; x <= ctx->lon_c; ++x) {
      const size_t ci = (size_t)(y - 1) * ctx->lon_c + (x - 1);
      real s = zero;
      for (int I = 0; I < 8; ++I) s = s + ctx->theta_s[8 * ci + I];
      if ((ctx->soil_tex[ci] > 0) && (ctx->soil_tex[ci] != 13) && (s > trunc_))
        ctx->land.push_back((int32_t)ci);
    }
  }
  ctx->last_diag.assign(ctx->land.size(), h9o_step_diag());
  ctx->have_soil = true;
  return 0;
}

int64_t h9o_num_land(const h9o_ctx* ctx) { return ctx ? (int64_t)ctx->land.size() : -1; }

int h9o_get_land_index(const h9o_ctx* ctx, int32_t* cell_xy) {
  if (!ctx || !cell_xy) return -1;
  std::copy(ctx->land.begin(), ctx->land.end(), cell_xy);
  return 0;
}

int h9o_init_state(h9o_ctx* ctx) { /* INIT.f90:707-811 */
  if (!ctx || !ctx->have_soil) return -1;
  const Geometry& g = ctx->geo;
  std::fill(ctx->h2osoi_liq.begin(), ctx->h2osoi_liq.end(), zero);
  std::fill(ctx->plant_mass.begin(), ctx->plant_mass.end(), zero);
  std::fill(ctx->zwt.begin(), ctx->zwt.end(), zero);
  std::fill(ctx->wa.begin(), ctx->wa.end(), zero);
  for (size_t k = 0; k < ctx->land.size(); ++k) {
    const size_t ci = (size_t)ctx->land[k];
    for (int I = 1; I <= 8; ++I) /* :730-731 */
      ctx->h2osoi_liq[8 * ci + I - 1] = K(0.4) * ctx->theta_s[8 * ci + I - 1] * g.dz[I] * rhow / K(1000.0);
    ctx->zwt[ci] = (g.zi[nlayers] + K(5000.0)) / K(1000.0); /* :739 */
    ctx->wa[ci] = K(4000.0);                                /* :744 */
    ctx->lai_litter[ci] = K(0.001);                         /* :748 */
    ctx->nplants[ci] = 1;                                   /* :752 */
    ctx->lai[ci] = zero;                                    /* :756 */
    for (int I = 0; I < 9; ++I) ctx->rootr_col[9 * ci + I] = zero; /* :760 */
    ctx->plant_mass[ci] = K(1.0);                                  /* :770 */
    ctx->plant_foliage_mass[ci] = K(0.0435);                       /* :771 */
    ctx->plant_length[ci] = rpow((K(400.0) * ctx->plant_mass[ci] / K(3.142E-3)), (one / K(3.0))); /* :776 */
    ctx->lai[ci] = ctx->lai[ci] + ctx->plant_foliage_mass[ci] * sla1 / plot_area; /* :781 */
    ctx->rdepth[ci] = K(0.3) * ctx->plant_length[ci];                             /* :786 */
    add_root_profile(g.zi, ctx->rdepth[ci], &ctx->rootr_col[9 * ci]);             /* :791-797 */
  }
  std::fill(ctx->smp.begin(), ctx->smp.end(), zero);
  return 0;
}

#define COPY_IN(dst, src, per)                                  \
  if (src) {                                                    \
    for (size_t k = 0; k < ctx->land.size(); ++k) {             \
      const size_t ci = (size_t)ctx->land[k];                   \
      for (int I = 0; I < (per); ++I) ctx->dst[(per)*ci + I] = src[(per)*ci + I]; \
    }                                                           \
  }
#define COPY_OUT(dst, src, per)                                 \
  if (dst) {                                                    \
    for (size_t k = 0; k < ctx->land.size(); ++k) {             \
      const size_t ci = (size_t)ctx->land[k];                   \
      for (int I = 0; I < (per); ++I) dst[(per)*ci + I] = ctx->src[(per)*ci + I]; \
    }                                                           \
  }

int h9o_set_state(h9o_ctx* ctx, const h9o_real* h2osoi_liq, const h9o_real* zwt,
                  const h9o_real* wa, const h9o_real* lai, const h9o_real* lai_litter,
                  const h9o_real* plant_mass, const h9o_real* plant_foliage_mass,
                  const h9o_real* plant_length, const h9o_real* rdepth, const h9o_real* rootr_col,
                  const int32_t* nplants, const h9o_real* smp) {
  if (!ctx || !ctx->have_soil) return -1;
  COPY_IN(h2osoi_liq, h2osoi_liq, 8)
  COPY_IN(zwt, zwt, 1)
  COPY_IN(wa, wa, 1)
  COPY_IN(lai, lai, 1)
  COPY_IN(lai_litter, lai_litter, 1)
  COPY_IN(plant_mass, plant_mass, 1)
  COPY_IN(plant_foliage_mass, plant_foliage_mass, 1)
  COPY_IN(plant_length, plant_length, 1)
  COPY_IN(rdepth, rdepth, 1)
  COPY_IN(rootr_col, rootr_col, 9)
  COPY_IN(nplants, nplants, 1)
  if (smp) {
    COPY_IN(smp, smp, 8)
  } else {
    std::fill(ctx->smp.begin(), ctx->smp.end(), zero);
  }
  for (int I = 0; I < 8; ++I) ctx->smp_shared[I] = zero;
  return 0;
}

int h9o_get_state(h9o_ctx* ctx, h9o_real* h2osoi_liq, h9o_real* zwt, h9o_real* wa, h9o_real* lai,
                  h9o_real* lai_litter, h9o_real* plant_mass, h9o_real* plant_foliage_mass,
                  h9o_real* plant_length, h9o_real* rdepth, h9o_real* rootr_col, int32_t* nplants,
                  h9o_real* smp) {
  if (!ctx || !ctx->have_soil) return -1;
  COPY_OUT(h2osoi_liq, h2osoi_liq, 8)
  COPY_OUT(zwt, zwt, 1)
  COPY_OUT(wa, wa, 1)
  COPY_OUT(lai, lai, 1)
  COPY_OUT(lai_litter, lai_litter, 1)
  COPY_OUT(plant_mass, plant_mass, 1)
  COPY_OUT(plant_foliage_mass, plant_foliage_mass, 1)
  COPY_OUT(plant_length, plant_length, 1)
  COPY_OUT(rdepth, rdepth, 1)
  COPY_OUT(rootr_col, rootr_col, 9)
  COPY_OUT(nplants, nplants, 1)
  COPY_OUT(smp, smp, 8)
  return 0;
}

int h9o_set_options(h9o_ctx* ctx, int loop_order, int smp_leak, int nthreads) {
  if (!ctx) return -1;
  if (smp_leak && (loop_order != 0 || nthreads > 1)) return -1; /* the leak is an order effect */
  ctx->loop_order = loop_order;
  ctx->smp_leak = smp_leak;
  ctx->nthreads = nthreads < 1 ? 1 : nthreads;
  return 0;
}

int h9o_set_real_evap(h9o_ctx* ctx, int on) {
  if (!ctx) return -1;
  ctx->real_evap = on ? 1 : 0;
  return 0;
}

int h9o_run_days(h9o_ctx* ctx, int ndays, const int32_t* year_index_of_day, const h9o_real* tas,
                 const h9o_real* rlds, const h9o_real* rsds, const h9o_real* huss,
                 const h9o_real* ps, const h9o_real* pr, const h9o_real* rhs) {
  if (!ctx || !ctx->have_soil || ndays < 0 || !year_index_of_day) return -1;
  if (!tas || !rlds || !rsds || !huss || !ps || !pr || !rhs) return -1;
  const size_t nl = ctx->land.size();
  const int day0 = (int)ctx->day_counter;
  if (ctx->loop_order == 1) { /* time outer, cells inner */
    for (int d = 0; d < ndays; ++d)
      for (size_t k = 0; k < nl; ++k) {
        const size_t ci = (size_t)ctx->land[k];
        cell_day(ctx, ci, forcing_at(ctx, ci, d, tas, rlds, rsds, huss, ps, pr, rhs),
                 year_index_of_day[d], day0 + d + 1);
      }
  } else { /* cells outer, time inner: HYBRID9.f90:120-295 */
    auto work = [&](size_t k0, size_t k1) {
      for (size_t k = k0; k < k1; ++k) {
        const size_t ci = (size_t)ctx->land[k];
        for (int d = 0; d < ndays; ++d)
          cell_day(ctx, ci, forcing_at(ctx, ci, d, tas, rlds, rsds, huss, ps, pr, rhs),
                   year_index_of_day[d], day0 + d + 1);
      }
    };
    const int nth = (int)std::min<size_t>((size_t)ctx->nthreads, std::max<size_t>(nl, 1));
    if (nth <= 1) {
      work(0, nl);
    } else {
      std::vector<std::thread> th;
      for (int t = 0; t < nth; ++t) th.emplace_back(work, nl * t / nth, nl * (t + 1) / nth);
      for (auto& t : th) t.join();
    }
  }
  ctx->day_counter += ndays;
  uint32_t any = 0;
  for (size_t k = 0; k < nl; ++k) any |= ctx->fault[ctx->land[k]];
  return (int)any;
}

int h9o_get_annual(h9o_ctx* ctx, int iyr, h9o_real* axy_npp, h9o_real* axy_plant_mass,
                   h9o_real* axy_rnf, h9o_real* axy_evap, h9o_real* axy_theta_total,
                   h9o_real* axy_theta) {
  if (!ctx || !ctx->have_soil || iyr < 1 || iyr > ctx->nyr) return -1;
  const size_t base = (size_t)(iyr - 1) * ctx->ncell;
  for (size_t k = 0; k < ctx->land.size(); ++k) {
    const size_t ci = (size_t)ctx->land[k];
    if (ctx->cur_year[ci] < iyr) continue; /* year not reached: leave the caller's fill */
    if (axy_npp) axy_npp[ci] = ctx->axy_npp[base + ci];
    if (axy_plant_mass) axy_plant_mass[ci] = ctx->axy_plant_mass[base + ci];
    if (axy_rnf) axy_rnf[ci] = ctx->axy_rnf[base + ci];
    if (axy_evap) axy_evap[ci] = ctx->axy_evap[base + ci];
    if (axy_theta_total) axy_theta_total[ci] = ctx->axy_theta_total[base + ci];
    if (axy_theta)
      for (int I = 0; I < 8; ++I) axy_theta[8 * ci + I] = ctx->axy_theta[8 * (base + ci) + I];
  }
  return 0;
}

int h9o_get_fault(h9o_ctx* ctx, uint32_t* any, uint32_t* code, int32_t* x, int32_t* y,
                  int32_t* day, int32_t* substep, h9o_real* imbalance, int64_t* n_faulted) {
  if (!ctx || !ctx->have_soil) return -1;
  uint32_t a = 0;
  int64_t nf = 0;
  long best = -1;
  for (size_t k = 0; k < ctx->land.size(); ++k) {
    const size_t ci = (size_t)ctx->land[k];
    if (!ctx->fault[ci]) continue;
    a |= ctx->fault[ci];
    ++nf;
    if (best < 0) {
      best = (long)ci;
    } else if (ctx->loop_order == 1) { /* earliest step, then lowest cell */
      const size_t b = (size_t)best;
      if (ctx->first_day[ci] < ctx->first_day[b] ||
          (ctx->first_day[ci] == ctx->first_day[b] && ctx->first_sub[ci] < ctx->first_sub[b]))
        best = (long)ci;
    }
  }
  if (any) *any = a;
  if (n_faulted) *n_faulted = nf;
  if (best >= 0) {
    const size_t b = (size_t)best;
    if (code) *code = ctx->first_code[b];
    if (x) *x = (int32_t)(b % ctx->lon_c) + 1;
    if (y) *y = (int32_t)(b / ctx->lon_c) + 1;
    if (day) *day = ctx->first_day[b];
    if (substep) *substep = ctx->first_sub[b];
    if (imbalance) *imbalance = ctx->first_imb[b];
  } else {
    if (code) *code = 0;
    if (x) *x = 0;
    if (y) *y = 0;
    if (day) *day = 0;
    if (substep) *substep = 0;
    if (imbalance) *imbalance = zero;
  }
  return 0;
}

int h9o_clear_fault(h9o_ctx* ctx) {
  if (!ctx) return -1;
  std::fill(ctx->fault.begin(), ctx->fault.end(), 0u);
  return 0;
}

int h9o_hydrology_step(h9o_ctx* ctx, const h9o_real* tas, const h9o_real* rlds,
                       const h9o_real* rsds, const h9o_real* huss, const h9o_real* ps,
                       const h9o_real* pr, const h9o_real* rhs, h9o_real* theta,
                       h9o_real* qflx_tran_veg_col, h9o_real* qflx_evap_grnd, h9o_real* rnf_inc,
                       h9o_real* w_imbalance, int32_t* jwt) {
  if (!ctx || !ctx->have_soil) return -1;
  if (!tas || !rlds || !rsds || !huss || !ps || !pr || !rhs) return -1;
  uint32_t any = 0;
  for (size_t k = 0; k < ctx->land.size(); ++k) {
    const size_t ci = (size_t)ctx->land[k];
    Cell c = ctx->cell(ci);
    Forcing f = forcing_at(ctx, ci, 0, tas, rlds, rsds, huss, ps, pr, rhs);
    Weather wx = derive_weather(f);
    h9o_step_diag& d = ctx->last_diag[k];
    hydrology(ctx->geo, c, f, wx, &d);
    note_fault(ctx, ci, d.fault, (int)ctx->day_counter + 1, 1, d.w1 - d.w0);
    any |= d.fault;
    if (theta)
      for (int I = 0; I < 8; ++I) theta[8 * ci + I] = d.theta[I];
    if (qflx_tran_veg_col) qflx_tran_veg_col[ci] = d.qflx_tran_veg_col;
    if (qflx_evap_grnd) qflx_evap_grnd[ci] = d.qflx_evap_grnd;
    if (rnf_inc) rnf_inc[ci] = d.rnf_inc;
    if (w_imbalance) w_imbalance[ci] = d.w1 - d.w0;
    if (jwt) jwt[ci] = d.jwt_final;
  }
  return (int)any;
}

int h9o_grow_day(h9o_ctx* ctx, const h9o_real* tas, h9o_real* npp, h9o_real* w_i, h9o_real* fT) {
  if (!ctx || !ctx->have_soil || !tas) return -1;
  for (size_t k = 0; k < ctx->land.size(); ++k) {
    const size_t ci = (size_t)ctx->land[k];
    Cell c = ctx->cell(ci);
    GrowDiag gd;
    grow(ctx->geo, c, tas[ci], &gd);
    if (npp) npp[ci] = gd.npp;
    if (w_i) w_i[ci] = gd.w_i;
    if (fT) fT[ci] = gd.fT;
  }
  return 0;
}

int h9o_last_step_diag(h9o_ctx* ctx, int x, int y, h9o_step_diag* out) {
  if (!ctx || !ctx->have_soil || !out) return -1;
  const int32_t ci = (int32_t)((size_t)(y - 1) * ctx->lon_c + (x - 1));
  auto it = std::lower_bound(ctx->land.begin(), ctx->land.end(), ci);
  if (it == ctx->land.end() || *it != ci) return -1;
  *out = ctx->last_diag[(size_t)(it - ctx->land.begin())];
  return 0;
}

int h9o_get_geometry(const h9o_ctx* ctx, h9o_real dz[10], h9o_real zc[10], h9o_real* dt) {
  if (!ctx || !ctx->configured) return -1;
  for (int I = 0; I < 10; ++I) {
    if (dz) dz[I] = ctx->geo.dz[I];
    if (zc) zc[I] = ctx->geo.zc[I];
  }
  if (dt) *dt = ctx->geo.dt;
  return 0;
}

int h9o_regrid_soil_layer(int lon_c, int lat_c, int layer, const h9o_real* theta_s_in,
                          const h9o_real* k_s_in, const h9o_real* lambda_in, const h9o_real* psi_s_in,
                          h9o_real* theta_s, h9o_real* hksat, h9o_real* bsw, h9o_real* psi_s) {
  if (lon_c < 1 || lat_c < 1 || layer < 1 || layer > 8) return -1;
  const size_t fw = (size_t)lon_c * 60;
  for (int y = 1; y <= lat_c; ++y)
    for (int x = 1; x <= lon_c; ++x) { /* INIT.f90:579-599 */
      real ts = zero, ks = zero, lm = zero, ps = zero;
      int j = 0;
      for (int x1 = (x - 1) * 60 + 1; x1 <= (x - 1) * 60 + 60; ++x1)
        for (int y1 = (y - 1) * 60 + 1; y1 <= (y - 1) * 60 + 60; ++y1) {
          const size_t k = (size_t)(y1 - 1) * fw + (size_t)(x1 - 1);
          if (theta_s_in[k] >= zero) {
            ts = ts + theta_s_in[k];
            ks = ks + k_s_in[k];
            lm = lm + lambda_in[k];
            ps = ps + psi_s_in[k];
            j = j + 1;
          }
        }
      if (j > 0) {
        ts = ts / (real)j;
        ks = ks / (real)j;
        lm = lm / (real)j;
        ps = ps / (real)j;
      }
      const size_t o = ((size_t)(y - 1) * lon_c + (x - 1)) * 8 + (layer - 1);
      theta_s[o] = ts / K(1.0E3);           /* :613 */
      hksat[o] = K(10.0) * ks / K(86400.0); /* :614 */
      real lambda = lm / K(1.0E3);          /* :615 */
      psi_s[o] = K(10.0) * ps;              /* :616 */
      lambda = rmax(lambda, trunc_);        /* :624 */
      bsw[o] = K(1.0) / lambda;             /* :628 */
    }
  return 0;
}

int h9o_time_boy(int year) { /* INIT.f90:844-859 */
  if (year < 1860 || year > 2300) return -1;
  int t = 1;
  for (int jyear = 1861; jyear <= year; ++jyear) {
    if ((jyear - 1) % 4 != 0) {
      t += 365;
    } else if ((jyear - 1) % 100 != 0) {
      t += 366;
    } else if ((jyear - 1) % 400 != 0) {
      t += 365;
    } else {
      t += 366;
    }
  }
  return t;
}

} /* extern "C" */
