/*
 * h9_physics_fast_tp.cuh -- the H9_MATH_FAST sub-step (HYDROLOGY.f90:141-1283) as the THROUGHPUT
 * kernel runs it: the 128-register build of days_kernel_fast, which steps a whole 0.5 deg (or
 * larger) grid in one wave with 14-16 warps per SM.
 *
 * Same arithmetic as hydrology_step_fast<kStepAllDeep / kStepGeneral> in h9_physics_fast.cuh --
 * the three give the same bits (tests/test_gpu_parity.py, tests/test_gpu_fullsize.py) -- but the
 * control flow of round 1: short branches around the rarely needed parts (the water-table search
 * and the Darcy recharge of cells with the table inside the column, the Drainage loops, the
 * cascade and repair), layers in pairs with their fluxes and Thomas rows behind them, the parts
 * of Recharge/Drainage that do not depend on the solve issued before it.  With 3.5 warps per
 * scheduler another warp fills a stall, and what counts is the number of executed instructions;
 * the straight-line layout that a lone warp needs (h9_physics_fast.cuh) executes more of them.
 * Measured on one B200, 0.5 deg, same box, ms per simulated year: 58.2 with this step, 59.4 with
 * the two-block layout in the same build (profiles/r02/README.md section 10).
 */
#ifndef H9_PHYSICS_FAST_TP_CUH
#define H9_PHYSICS_FAST_TP_CUH

#include "h9_physics_fast.cuh"

namespace h9 {

template <class C>
__device__ __forceinline__ uint32_t hydrology_step_fast_tp(const Geo& g, const C& c, const DayFast& d,
                                                        FastState& s, StepOut& o) {
  constexpr float kLog2e = 1.4426950408889634f;
  uint32_t fault = 0;
  const float dt = g.dt, rdt = g.rdt;
  float theta[NL];

  /* :141-151 */
#pragma unroll
  for (int i = 0; i < NL; ++i) theta[i] = s.h2o[i] * g.rdzw[i + 1];
  /* column sums as a pairwise tree: 3 dependent adds instead of 8 */
  const float w0 = (d.rain_dt + s.wa) + (((s.h2o[0] + s.h2o[1]) + (s.h2o[2] + s.h2o[3])) +
                                         ((s.h2o[4] + s.h2o[5]) + (s.h2o[6] + s.h2o[7])));

  /* SurfaceRunoff :182-212 */
  const float fsat = c.fmax() * MathFast::ex2((-0.5f * kFff * kLog2e) * s.zwt);
  float qflx_surf = fsat * d.forc_rain;

  /* beta from the previous sub-step's smp :269-276: 1 - x/(-150000) == 1 + x/150000 */
  float bw[NL];
#pragma unroll
  for (int i = 0; i < NL; ++i)
    bw[i] = __saturatef(fmaf(s.smp[i], 1.0f / 150000.0f, g.kbeta[i + 1]));
  const float4 ra = c.rootr4(0), rb = c.rootr4(1);
  const float rr[NL] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
  const float beta = fmaf(rr[0], bw[0], fmaf(rr[1], bw[1], fmaf(rr[2], bw[2], rr[3] * bw[3]))) +
                     fmaf(rr[4], bw[4], fmaf(rr[5], bw[5], fmaf(rr[6], bw[6], rr[7] * bw[7])));

  /* rsc :283-295, rss :325-331 */
  float rsc = (d.canopy_on && beta > 0.0f) ? d.rsc_num * MathFast::rcp(d.rsc_den0 * beta) : 1.0E6f;
  rsc = fmaxf(rsc, d.rsc_floor);
  const float rss = (theta[0] <= 0.15f)
                        ? d.litter10 * MathFast::ex2((35.63f * kLog2e) * (0.15f - theta[0]))
                        : fmaf(d.litter1000, 1.0f - theta[0] * c.inv_ths(0), 10.0f);

  /* two-source Penman-Monteith :344-389 */
  /* two reciprocals from one MUFU: 1/a = b/(a*b), 1/b = a/(a*b) (both denominators are
   * positive and of order 1e-4 .. 1e2) */
  const float dPMc = fmaf(d.gamma, fmaf(rsc, d.inv_raa_rac, 1.0f), d.desatdT);
  const float dPMs = fmaf(d.gamma, fmaf(rss, d.inv_raa_ras, 1.0f), d.desatdT);
  const float Rs = fmaf(d.gamma, rss, d.dg_ras);
  const float Rc = fmaf(d.gamma, rsc, d.dg_rac);
  /* Cc = 1/(1+Rc*Ra/(Rs*(Rc+Ra))) and Cs = 1/(1+Rs*Ra/(Rc*(Rs+Ra))) over the common
   * denominator Rs*Rc + Rs*Ra + Rc*Ra */
  const float RsRc = Rs * Rc, RsRa = Rs * d.Ra, RcRa = Rc * d.Ra;
  /* three reciprocals from one MUFU */
  const float dD = RsRc + RsRa + RcRa;
  const float pPM = dPMc * dPMs;
  const float r3 = MathFast::rcp(pPM * dD);
  const float invD = r3 * pPM;
  const float rPM = r3 * dD;
  const float PMc = d.pmc_num * (rPM * dPMs);
  const float PMs = d.pms_num * (rPM * dPMc);
  const float LE = fmaf(RsRc + RsRa, PMc, (RsRc + RcRa) * PMs) * invD;
  const float VDD0 = fmaf(d.A - d.dg * LE, d.raa_inv_rhocp, d.VDD);
  const float rv = d.rhocp * VDD0;
  const float dLEc = fmaf(d.gamma, fmaf(rsc, d.inv_rac, 1.0f), d.desatdT);
  const float dLEs = fmaf(d.gamma, fmaf(rss, d.inv_ras, 1.0f), d.desatdT);
  const float rLE = MathFast::rcp(dLEc * dLEs);
  const float LEc = fmaf(rv, d.inv_rac, d.lec_a) * (rLE * dLEs);
  const float LEs = fmaf(rv, d.inv_ras, d.les_a) * (rLE * dLEc);
  const float qflx_tran_veg_col = LEc * d.k_lamb;
  float qflx_evap_grnd = LEs * d.k_lamb;

  /* evaporation limit :396-400 */
  const float evap_max1 =
      fmaxf(0.0f, fmaf(-qflx_tran_veg_col, rr[0], g.dz[1] * (theta[0] - kWatmin) * rdt));
  qflx_evap_grnd = fminf(evap_max1, qflx_evap_grnd);

  /* Infiltration :426-478 */
  const float qflx_in_soil = (d.forc_rain - qflx_surf) - qflx_evap_grnd;
  const float qinmax = (1.0f - fsat) * c.hksmin3();
  const float qflx_infl_excess = fmaxf(0.0f, qflx_in_soil - qinmax);
  const float qflx_infl = qflx_in_soil - qflx_infl_excess;
  qflx_surf += qflx_infl_excess;

  /* SoilWater :492-508 */
  float zwtmm = 1000.0f * s.zwt;
  /* jwt == NL  <=>  zwt > zi(8)/1000 (the interfaces increase); the full search only runs for
   * cells whose water table is inside the column */
  const bool deep = !(s.zwt <= g.zim[NL]);
  int jwt = NL;
  if (!deep) jwt = find_jwt(g, s.zwt);

  /* equilibrium profile :517-590, branch-free.  With B0 = (-psi+zwtmm-zi(I-1))/(-psi) and
   * BI = (-psi+zwtmm-zi(I))/(-psi), the three cases of the reference are one formula:
   *   vol = [ A*(max(BI,1)**e1 - max(B0,1)**e1) + ths*max(zi(I)-zwtmm,0) ] / dz,  A = psi*ths/e1
   * below the layer both bases are >= 1 and the last term vanishes (:548-558); inside it
   * BI < 1 so the first power is tempi = 1 (:532-540); above it both powers are 1 and
   * (zi(I)-zwtmm)/dz >= 1 so the clamp to theta_s returns theta_s (:523). */
  /* Layers are processed in pairs and each pair is followed at once by the interface fluxes it
   * completes and by the rows of the Thomas forward sweep those fluxes complete: the sweep's
   * dependent chain (one reciprocal per row) then runs behind the powers of the deeper layers
   * instead of after all of them.  Same operations on the same operands as the separate
   * loops: the results are bit-identical. */
  float zq[NL + 1];
  float hk[NL], dhkdw[NL], dsmpdw[NL], snode[NL];
  float q[NL], qa[NL], qb[NL]; /* per interface i (below layer i): flux and its two derivatives */
  float dwat2[NL + 1], gam[NL + 1];
  float bet = 1.0f, rbet = 1.0f, minpiv = 3.0e38f;
  const float dz9 = deep ? (zwtmm - g.zc[NL]) : g.dz[NL];
  const float4 rc = c.rootr4(0), rd = c.rootr4(1); /* re-read: cheaper than eight live registers */
  const float rs[NL] = {rc.x, rc.y, rc.z, rc.w, rd.x, rd.y, rd.z, rd.w};
  /* parts of Recharge/Drainage (:856-965) that do not depend on the solve: the aquifer's
   * specific yield and its reciprocal */
  const float rous_early = fast_specific_yield(c, NL - 1, zwtmm);
  const float rrous_early = MathFast::rcp(rous_early);

  auto eq_layer = [&](int i) { /* equilibrium profile :517-573 of layer i+1 */
    const float zlo = g.zi[i], zhi = g.zi[i + 1];
    const float4 A = c.g0(i), B = c.g1(i); /* two 128-bit shared-memory loads per layer */
    const float psi = A.x, inv_npsi = A.y, e1 = A.z, ths = B.x;
    const float t0 = fast_pow(fmaxf(fmaf(zwtmm - zlo, inv_npsi, 1.0f), 1.0f), e1);
    const float ti = fast_pow(fmaxf(fmaf(zwtmm - zhi, inv_npsi, 1.0f), 1.0f), e1);
    float vol = fmaf(A.w, ti - t0, ths * fmaxf(fmaf(-zwtmm, g.rdzl[i + 1], g.zhr[i + 1]), 0.0f));
    vol = fminf(ths, fmaxf(vol, 0.0f));
    zq[i] = fmaxf(kSmpmin, psi * fast_pow(fmaxf(vol * B.y, 0.01f), -B.z));
  };
  auto hk_layer = [&](int i) { /* hk, dhkdw, smp, dsmpdw :598-639 of layer i+1 */
    const int ip = (i + 1 < NL) ? i + 1 : NL - 1;
    const float4 B = c.g1(i), G = c.g2(i);
    const float b = B.z;
    const float s1 = fminf(1.0f, (theta[i] + theta[ip]) * G.x);
    const float s2 = B.w * fast_pow(s1, fmaf(2.0f, b, 2.0f));
    hk[i] = s1 * s2;
    dhkdw[i] = G.z * s2; /* (2b+3)*s2/(ths(I)+ths(I+1)) */
    const float s_node = fminf(1.0f, fmaxf(theta[i] * B.y, 0.01f));
    const float sm = fmaxf(kSmpmin, G.y * fast_pow(s_node, -b));
    s.smp[i] = sm;
    snode[i] = s_node;
    dsmpdw[i] = G.w * sm; /* (-b/ths)*smp, still to be divided by s_node */
  };
  auto pair_rcp = [&](int k) { /* two reciprocals from one MUFU; s_node is in [0.01, 1] */
    const float r = MathFast::rcp(snode[k] * snode[k + 1]);
    dsmpdw[k] *= r * snode[k + 1];
    dsmpdw[k + 1] *= r * snode[k];
  };
  auto flux = [&](int i) { /* interface below layer i+1, i = 0..6 */
    const float rden = g.rden[i + 1];
    const float num = (s.smp[i + 1] - s.smp[i]) - (zq[i + 1] - zq[i]);
    const float nd = num * dhkdw[i];
    q[i] = -hk[i] * num * rden;                     /* qout(I) == qin(I+1) */
    qa[i] = fmaf(hk[i], dsmpdw[i], -nd) * rden;     /* dqodw1(I) == dqidw0(I+1) */
    qb[i] = -fmaf(hk[i], dsmpdw[i + 1], nd) * rden; /* dqodw2(I) == dqidw1(I+1) */
  };
  auto row0 = [&]() {
    const float rmx = fmaf(-qflx_tran_veg_col, rs[0], qflx_infl - q[0]);
    bet = g.dzdt[1] + qa[0];
    if (bet == 0.0f) fault |= FAULT_PIVOT1;
    rbet = MathFast::rcp(bet);
    dwat2[0] = rmx * rbet;
  };
  auto row = [&](int i) { /* rows 2..8 (i = 1..7) */
    const float rmx = fmaf(-qflx_tran_veg_col, rs[i], q[i - 1] - q[i]);
    const float amx = -qa[i - 1];
    const float bmx = (g.dzdt[i + 1] - qb[i - 1]) + qa[i];
    gam[i] = qb[i - 1] * rbet; /* cmx(I-1)/BET */
    bet = fmaf(-amx, gam[i], bmx);
    minpiv = fminf(minpiv, fabsf(bet));
    rbet = MathFast::rcp(bet);
    dwat2[i] = fmaf(-amx, dwat2[i - 1], rmx) * rbet;
  };

  { /* :576-590: vol_eq(9), evaluated for every cell and selected by `deep` (see below) */
    const int i = NL - 1;
    const float ths = c.ths(i), psi = c.psi(i), e1 = c.e1(i), inv_npsi = c.inv_npsi(i);
    const float u = (zwtmm - g.zi[NL]) * inv_npsi;
    const float a2 = 0.5f * (e1 - 1.0f), a3 = (1.0f / 3.0f) * (e1 - 2.0f), a4 = 0.25f * (e1 - 3.0f);
    const float r_series = e1 * fmaf(u * a2, fmaf(u * a3, fmaf(u, a4, 1.0f), 1.0f), 1.0f);
    const float r_direct = (fast_pow(1.0f + u, e1) - 1.0f) * MathFast::rcp(u);
    const float r = (u < 0.03f) ? r_series : r_direct;
    const float coefA = c.coef3(i) * (g.zi[NL] - g.zi[NL - 1]);
    float vol = -coefA * inv_npsi * r;
    vol = fminf(ths, fmaxf(vol, 0.0f));
    const float z9 = fmaxf(kSmpmin, psi * fast_pow(fmaxf(vol * c.inv_ths(i), 0.01f), -c.bsw(i)));
    zq[NL] = deep ? z9 : 0.0f;
  }
#pragma unroll
  for (int P = 0; P < NL / 2; ++P) {
    eq_layer(2 * P);
    eq_layer(2 * P + 1);
    hk_layer(2 * P);
    hk_layer(2 * P + 1);
    pair_rcp(2 * P);
    if (P > 0) flux(2 * P - 1);
    flux(2 * P);
    if (P == 0) {
      row0();
    } else {
      row(2 * P - 1);
      row(2 * P);
    }
  }
  { /* aquifer node :645-650 and its interface :737-753; inert when the table is in the column */
    const int i = NL - 1;
    const float b = c.bsw(i), inv_ths = c.inv_ths(i);
    const float s_node = fminf(1.0f, fmaxf(0.5f * fmaf(theta[i], inv_ths, 1.0f), 0.01f));
    const float smp1 = fmaxf(kSmpmin, c.psi(i) * fast_pow(s_node, -b));
    const float zc9 = 0.5f * (zwtmm + g.zc[NL]);
    const float den9 = zc9 - g.zc[NL];
    const float r9 = MathFast::rcp(s_node * den9);
    const float rden = r9 * s_node;
    const float dsmpdw1 = c.comp(2, i, 3) * smp1 * (r9 * den9);
    const float num = (smp1 - s.smp[i]) - (zq[NL] - zq[i]);
    const float nd = num * dhkdw[i];
    q[i] = deep ? -hk[i] * num * rden : 0.0f;
    qa[i] = deep ? fmaf(hk[i], dsmpdw[i], -nd) * rden : 0.0f;
    qb[i] = deep ? -fmaf(hk[i], dsmpdw1, nd) * rden : 0.0f;
  }
  row(NL - 1);
  { /* aquifer row */
    const float rmx = q[NL - 1];
    const float amx = -qa[NL - 1];
    const float bmx = fmaf(dz9, rdt, -qb[NL - 1]);
    gam[NL] = qb[NL - 1] * rbet;
    bet = fmaf(-amx, gam[NL], bmx);
    minpiv = fminf(minpiv, fabsf(bet));
    dwat2[NL] = fmaf(-amx, dwat2[NL - 1], rmx) * MathFast::rcp(bet);
  }
  /* for a water table inside the column the Darcy recharge and the specific yield of the
   * layer holding the table do not depend on the solve either (issued before the back
   * substitution; they read the NEW smp, :880) */
  float qcharge_early = 0.0f, sy_first = 0.02f;
  if (!deep) {
    const int jm = (jwt > 1 ? jwt : 1) - 1;
    const float th_j = pick<NL>(theta, jwt);
    const float s1 = fminf(1.0f, fmaxf(th_j * c.inv_ths(jwt), 0.01f));
    const float ka = c.hks(jwt) * fast_pow(s1, fmaf(2.0f, c.bsw(jwt), 3.0f));
    const float smp1 = fmaxf(kSmpmin, pick<NL>(s.smp, jm));
    float zq_j = zq[0];
#pragma unroll
    for (int k = 1; k < NL; ++k)
      if (jm == k) zq_j = zq[k];
    const float wh = smp1 - zq_j;
    const float denom = (jwt == 0) ? (zwtmm + 1.0f) : (zwtmm - g.zc[jwt]) * 2.0f;
    qcharge_early = ka * wh * MathFast::rcp(denom); /* -ka*(0 - wh)/denom */
    qcharge_early = fminf(g.q10_hi, fmaxf(g.q10_lo, qcharge_early));
    sy_first = fast_specific_yield(c, jwt, zwtmm); /* layer jwt+1 */
  }
  if (minpiv == 0.0f) fault |= FAULT_PIVOT2; /* BET == 0 in some row :818 */
  /* recharge :856-904 */
  float qcharge;
  qcharge = deep ? dwat2[NL] * dz9 * rdt : qcharge_early;
  /* Drainage :923-1009.  The jwt of :923-931 equals the one of :499-508: zwt has not
   * changed in between */
  float rous = rous_early;
  const int jfirst = jwt + 1;
  if (jwt == NL) {
    s.wa = fmaf(qcharge, dt, s.wa);
    s.zwt = fmaf(-(qcharge * dt * 0.001f), rrous_early, s.zwt);
  } else { /* zwtmm stays the stale value of :492 inside the loops */
    float qcharge_tot = qcharge * dt;
    if (qcharge_tot > 0.0f) { /* rising, layers jwt+1 .. 1 */
      for (int I = jwt + 1; I >= 1; --I) {
        const float s_y = (I == jfirst) ? sy_first : fast_specific_yield(c, I - 1, zwtmm);
        const float ql = fmaxf(fminf(qcharge_tot, s_y * (zwtmm - g.zi[I - 1])), 0.0f);
        s.zwt -= ql * MathFast::rcp(s_y) * 0.001f; /* s_y >= 0.02 > 0 */
        qcharge_tot -= ql;
        if (qcharge_tot <= 0.0f) break;
      }
    } else { /* deepening, layers jwt+1 .. 8 */
      for (int I = jwt + 1; I <= NL; ++I) {
        const float s_y = (I == jfirst) ? sy_first : fast_specific_yield(c, I - 1, zwtmm);
        const float ql = fminf(fmaxf(qcharge_tot, -s_y * (g.zi[I] - zwtmm)), 0.0f);
        qcharge_tot -= ql;
        if (qcharge_tot >= 0.0f) {
          s.zwt -= ql * MathFast::rcp(s_y) * 0.001f;
          break;
        }
        s.zwt = g.zim[I];
      }
      if (qcharge_tot > 0.0f) s.zwt -= qcharge_tot * 0.001f * rrous_early;
    }
    jwt = find_jwt(g, s.zwt);
  }

  zwtmm = 1000.0f * s.zwt; /* :1015 */

  /* baseflow :1024-1118 */
  float rsub_top = 5.5E-3f * MathFast::ex2((-kFff * kLog2e) * s.zwt);
  rous = fast_specific_yield(c, NL - 1, zwtmm);
  /* dwat2(9), the only unknown Recharge and Drainage read, is final after the forward sweep;
   * the back substitution :828-830 and the update of h2osoi_liq :845-850 are issued here,
   * behind the water-table update and the baseflow's exp / pow, just before baseflow
   * touches h2osoi_liq */
#pragma unroll
  for (int i = NL - 1; i >= 0; --i) dwat2[i] = fmaf(-gam[i + 1], dwat2[i + 1], dwat2[i]);
#pragma unroll
  for (int i = 0; i < NL; ++i) s.h2o[i] = fmaf(dwat2[i], g.dz[i + 1], s.h2o[i]);
  if (jwt == NL) { /* jwt is not recomputed on this path */
    s.wa = fmaf(-rsub_top, dt, s.wa);
    s.zwt = fmaf(rsub_top * dt * 0.001f, MathFast::rcp(rous), s.zwt);
    s.h2o[NL - 1] += fmaxf(0.0f, s.wa - 5000.0f);
    s.wa = fminf(s.wa, 5000.0f);
  } else {
    float rsub_top_tot = -rsub_top * dt;
    if (rsub_top_tot > 0.0f) {
      fault |= FAULT_RSUB;
    } else {
      for (int I = jwt + 1; I <= NL; ++I) {
        const float s_y = fast_specific_yield(c, I - 1, zwtmm);
        const float rl = fminf(fmaxf(rsub_top_tot, -(s_y * (g.zi[I] - zwtmm))), 0.0f);
#pragma unroll
        for (int k = 0; k < NL; ++k)
          s.h2o[k] = (k == I - 1) ? s.h2o[k] + rl : s.h2o[k]; /* a select, not a jump table */
        rsub_top_tot -= rl;
        if (rsub_top_tot >= 0.0f) {
          s.zwt -= rl * MathFast::rcp(s_y) * 0.001f;
          break;
        }
        s.zwt = g.zim[I];
      }
      s.zwt -= rsub_top_tot * 0.001f * MathFast::rcp(rous); /* residual, unconditional :1100-1101 */
      s.wa += rsub_top_tot;
    }
    jwt = -1; /* :1110-1116, evaluated from zwt_j only where it is read (dryness repair, output) */
  }
  const float zwt_j = s.zwt;

  s.zwt = fminf(80.0f, fmaxf(0.0f, s.zwt)); /* :1122-1123 */

  /* excess cascade :1131-1152 and dryness repair :1161-1205: both are no-ops unless some layer
   * is above its capacity / below watmin, so ONE rarely taken branch guards them.  The repair's
   * trigger is taken before the cascade: the cascade only lowers layers to their capacity
   * (>= 0.01*dz > watmin) and raises others, so it cannot create a layer below watmin, and
   * inside the branch the repair works from the current values anyway. */
  float over = s.h2o[0] - fmaxf(0.0f, c.ths(0) * g.dz[1]);
  float lowest = s.h2o[0];
#pragma unroll
  for (int i = 1; i < NL; ++i) {
    over = fmaxf(over, fmaf(-fmaxf(0.01f, c.ths(i)), g.dz[i + 1], s.h2o[i]));
    lowest = fminf(lowest, s.h2o[i]);
  }
  float qflx_rsub_sat = 0.0f;
  float xs = 0.0f;
  if (over > 0.0f || lowest < kWatmin) {
#pragma unroll
    for (int i = NL - 1; i >= 1; --i) {
      const float cap = fmaxf(0.01f, c.ths(i)) * g.dz[i + 1];
      const float xsi = fmaxf(s.h2o[i] - cap, 0.0f);
      s.h2o[i] = fminf(cap, s.h2o[i]);
      s.h2o[i - 1] += xsi;
    }
    const float cap1 = fmaxf(0.0f, c.ths(0) * g.dz[1]);
    const float xs1 = fmaxf(fmaxf(s.h2o[0], 0.0f) - cap1, 0.0f);
    s.h2o[0] = fminf(cap1, s.h2o[0]);
    qflx_rsub_sat = xs1 * rdt;

    if (jwt < 0) jwt = find_jwt(g, zwt_j);
#pragma unroll
    for (int i = 0; i < NL - 1; ++i) {
      xs = fmaxf(kWatmin - s.h2o[i], 0.0f); /* > 0 exactly when h2o < watmin */
      if (xs > 0.0f && i + 1 == jwt) s.zwt += xs * MathFast::rcp(fmaxf(0.01f, c.ths(i))) * 0.001f;
      s.h2o[i] += xs;
      s.h2o[i + 1] -= xs;
    }
    xs = 0.0f;
    if (s.h2o[NL - 1] < kWatmin) { /* search upward for water :1181-1198 */
      xs = kWatmin - s.h2o[NL - 1];
      bool done = false;
#pragma unroll
      for (int j = NL - 2; j >= 0; --j) {
        if (!done) {
          const float avail = fmaxf(s.h2o[j] - kWatmin - xs, 0.0f);
          const float take = (avail >= xs) ? xs : avail;
          done = (avail >= xs);
          s.h2o[NL - 1] += take;
          s.h2o[j] -= take;
          xs = done ? 0.0f : xs - take;
        }
      }
    }
  }
  s.h2o[NL - 1] += xs;  /* :1205 */
  rsub_top -= xs * rdt; /* :1211 */

  /* balance :1221-1244 */
  const float w1 =
      fmaf((qflx_surf + qflx_evap_grnd + qflx_tran_veg_col) + rsub_top + qflx_rsub_sat, dt, s.wa) +
      (((s.h2o[0] + s.h2o[1]) + (s.h2o[2] + s.h2o[3])) + ((s.h2o[4] + s.h2o[5]) + (s.h2o[6] + s.h2o[7])));
  const float imb = w1 - w0;
  if (!(fabsf(imb) <= 0.1f)) fault |= FAULT_IMBAL;

  /* :1282-1283 */
  const float r1 = qflx_surf * dt, r2 = rsub_top * dt;
  s.rnf_sum = fmaf(rsub_top, dt, fmaf(qflx_surf, dt, s.rnf_sum));

  o.qflx_tran_veg_col = qflx_tran_veg_col;
  o.qflx_evap_grnd = qflx_evap_grnd;
  o.rnf_inc = r1 + r2;
  o.imbalance = imb;
  o.jwt = (jwt < 0) ? find_jwt(g, zwt_j) : jwt; /* dead code where the caller ignores it (K3) */
  return fault;
}

} /* namespace h9 */
#endif
