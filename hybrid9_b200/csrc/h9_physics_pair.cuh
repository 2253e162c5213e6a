/*
 * h9_physics_pair.cuh -- the H9_MATH_FAST sub-step (HYDROLOGY.f90:141-1283) with TWO lanes per
 * land cell, for shards so small that the thread-per-cell kernel leaves most schedulers
 * empty (one latitude band of a 0.5 deg grid on 8 GPUs is 8.4k cells = 264 warps on 592
 * schedulers: the time is one warp's instruction stream, whatever the cell count).
 *
 * Mapping.  Lanes 2p and 2p+1 of a warp share one cell.  The even lane (h = 0) owns soil
 * layers 1..4, the odd lane (h = 1) owns layers 8..5 IN REVERSE ORDER plus the aquifer node:
 *
 *      h = 0:  surface | L1  L2  L3  L4 |            local j = 0..3  <->  layer j+1
 *      h = 1:  aquifer | L8  L7  L6  L5 |            local j = 0..3  <->  layer 8-j
 *                  ^ outer node (-1)    ^ junction (between L4 and L5)
 *
 * Both lanes therefore run the SAME instruction stream "outer boundary -> four layers ->
 * junction": per-layer work (equilibrium profile :517-573, matric potential :626-639) on
 * four layers each, interface conductivities and Darcy fluxes (:598-621, :661-753) on the
 * five interfaces k = -1..3 between local nodes k and k+1, written as the INWARD flux
 *      F(k) = -hk(k) * ((smp(k+1)-smp(k)) - (zq(k+1)-zq(k))) / (zc(k+1)-zc(k))
 * which is q(I) of the reference for h = 0 and -q(I) for h = 1, with hk(k) built from the
 * UPPER layer's hksat/bsw as the reference does (:605-621; the constants are tabulated per
 * interface so the code does not branch on h).  The tridiagonal system (:661-799) is the
 * same nine equations; it is solved from both ends at once (lane 0 eliminates downward from
 * row 1, lane 1 upward from the aquifer row 9) and closed by the 2x2 system at the junction,
 * instead of :806-831's single forward sweep: five dependent pivots instead of nine, no
 * lane-to-lane hand-off.  Fast mode is tolerance-gated (not bit-gated) against the oracle;
 * the operation order of the solve and of the column sums differs from the thread-per-cell
 * kernel, everything else is the same arithmetic.
 *
 * Per-cell scalars (energy balance, water table, recharge, drainage, balance) are computed
 * by both lanes (the warp issues the instruction once either way).  Cross-lane traffic per
 * sub-step: 13 shuffles (two boundary thetas, three column sums, three junction node values,
 * two junction unknowns, dwat2(9), the fault/repair flags).
 */
#ifndef H9_PHYSICS_PAIR_CUH
#define H9_PHYSICS_PAIR_CUH

#include "h9_physics_fast.cuh"

namespace h9 {

constexpr int NH = NL / 2; /* layers per lane */
constexpr unsigned kFullMask = 0xffffffffu;

/* table rows (float4 units) of one lane's shared-memory column */
constexpr int kPairRowG0 = 0;            /* [NH] (psi, 1/(-psi), 1-1/b, psi*ths/(1-1/b)/dz)   */
constexpr int kPairRowG1 = NH;           /* [NH] (ths, 1/ths, b, -b/ths)                       */
constexpr int kPairRowGI = 2 * NH;       /* [NH+1] interface k=-1..3: (its, hks, 2b+2, (2b+3)*its) */
constexpr int kPairRowCap = 3 * NH + 1;  /* capacities of the four layers :1131-1148           */
constexpr int kPairRowRoot = 3 * NH + 2; /* rootr_col of the four layers                       */
constexpr int kPairRowMisc = 3 * NH + 3; /* (Fmax, MINVAL(hksat(1:3)), rootr_col(1), 1/ths(1)) */
constexpr int kPairRowF = 3 * NH + 4;    /* [NL] whole column: (ths, 1/(-psi), -1/b, 1/ths)    */
constexpr int kPairRowF2 = kPairRowF + NL; /* [NL] whole column: (hksat, b, -, -)              */
constexpr int kPairRows = kPairRowF2 + NL;
constexpr int kPairFloatsPerLane = kPairRows * 4;

/* geometry of the lane's four layers and interfaces, in registers (a lane-dependent index
 * into the constant bank would serialise the constant cache) */
struct PairGeo {
  float zlo[NH], zhi[NH], dz[NH], rdzw[NH], dzdt[NH], rdzl[NH], zhr[NH], kbeta[NH];
  float rden[NH]; /* 1/(zc(k+1)-zc(k)) of interface k = 0..3 (3 = junction) */
  __device__ __forceinline__ void init(const Geo& g, int h) {
#pragma unroll
    for (int j = 0; j < NH; ++j) {
      const int I = h ? NL - j : j + 1; /* 1-based layer */
      zlo[j] = g.zi[I - 1];
      zhi[j] = g.zi[I];
      dz[j] = g.dz[I];
      rdzw[j] = g.rdzw[I];
      dzdt[j] = g.dzdt[I];
      rdzl[j] = g.rdzl[I];
      zhr[j] = g.zhr[I];
      kbeta[j] = g.kbeta[I];
      /* interface j lies below layer j+1 (h = 0) / below layer 7-j (h = 1) */
      rden[j] = h ? g.rden[NL - 1 - j] : g.rden[j + 1];
    }
  }
};

template <int STRIDE>
struct PairTable {
  float4* base;
  __device__ __forceinline__ float4 row(int r) const { return base[r * STRIDE]; }
  __device__ __forceinline__ float4 g0(int j) const { return base[(kPairRowG0 + j) * STRIDE]; }
  __device__ __forceinline__ float4 g1(int j) const { return base[(kPairRowG1 + j) * STRIDE]; }
  __device__ __forceinline__ float4 gi(int k) const { return base[(kPairRowGI + 1 + k) * STRIDE]; }
  __device__ __forceinline__ float4 cap() const { return base[kPairRowCap * STRIDE]; }
  __device__ __forceinline__ float4 rootr4() const { return base[kPairRowRoot * STRIDE]; }
  __device__ __forceinline__ float4 misc() const { return base[kPairRowMisc * STRIDE]; }
  __device__ __forceinline__ float comp(int r, int k) const {
    H9_ASSERT(r >= 0 && r < kPairRows && k >= 0 && k < 4);
    return reinterpret_cast<const float*>(&base[r * STRIDE])[k];
  }
  /* whole-column accessors used by the data-dependent Drainage code (dynamic layer index) */
  __device__ __forceinline__ float ths(int i) const { return comp(kPairRowF + i, 0); }
  __device__ __forceinline__ float inv_npsi(int i) const { return comp(kPairRowF + i, 1); }
  __device__ __forceinline__ float e1(int i) const { return comp(kPairRowF + i, 2) + 1.0f; }
  __device__ __forceinline__ float me(int i) const { return comp(kPairRowF + i, 2); } /* -1/b */
  __device__ __forceinline__ float inv_ths(int i) const { return comp(kPairRowF + i, 3); }
  __device__ __forceinline__ float hks(int i) const { return comp(kPairRowF2 + i, 0); }
  __device__ __forceinline__ float bsw(int i) const { return comp(kPairRowF2 + i, 1); }

  __device__ __forceinline__ void set_rootr(const float (&rootr)[NL], int h) const {
    base[kPairRowRoot * STRIDE] = h ? make_float4(rootr[7], rootr[6], rootr[5], rootr[4])
                                    : make_float4(rootr[0], rootr[1], rootr[2], rootr[3]);
    reinterpret_cast<float*>(&base[kPairRowMisc * STRIDE])[2] = rootr[0];
  }

  __device__ __forceinline__ void init(const Geo& g, const Params& p, const float (&rootr)[NL], int h) const {
    const float hmin = fminf(fminf(p.hksat[0], p.hksat[1]), p.hksat[2]);
    float capv[NH];
#pragma unroll
    for (int j = 0; j < NH; ++j) {
      const int i = h ? NL - 1 - j : j; /* 0-based layer */
      const float psi = p.psi_s[i], ths = p.theta_s[i], b = p.bsw[i];
      const float e1 = 1.0f - 1.0f / b;
      base[(kPairRowG0 + j) * STRIDE] =
          make_float4(psi, 1.0f / (-psi), e1, psi * ths / e1 / (g.zi[i + 1] - g.zi[i]));
      base[(kPairRowG1 + j) * STRIDE] = make_float4(ths, 1.0f / ths, b, -b / ths);
      /* :1131-1148: the top layer spills above theta_s*dz, the others above eff_porosity*dz */
      capv[j] = (i == 0) ? fmaxf(0.0f, ths * g.dz[1]) : fmaxf(0.01f, ths) * g.dz[i + 1];
    }
    base[kPairRowCap * STRIDE] = make_float4(capv[0], capv[1], capv[2], capv[3]);
#pragma unroll
    for (int k = -1; k < NH; ++k) {
      /* the layer above interface k: h = 0 -> layer k (0-based; k = -1 is the surface, a dummy),
       * h = 1 -> layer 6-k (k = -1: layer 7, whose lower neighbour is itself, :605) */
      int u = h ? NL - 2 - k : k;
      if (u < 0) u = 0;
      const int l = (u + 1 < NL) ? u + 1 : NL - 1;
      const float its = 1.0f / (p.theta_s[u] + p.theta_s[l]);
      const float b = p.bsw[u];
      base[(kPairRowGI + 1 + k) * STRIDE] =
          make_float4(its, p.hksat[u], fmaf(2.0f, b, 2.0f), (2.0f * b + 3.0f) * its);
    }
    base[kPairRowMisc * STRIDE] = make_float4(p.fmax, hmin, rootr[0], 1.0f / p.theta_s[0]);
#pragma unroll
    for (int i = 0; i < NL; ++i) {
      const float b = p.bsw[i];
      const float e1 = 1.0f - 1.0f / b;
      base[(kPairRowF + i) * STRIDE] =
          make_float4(p.theta_s[i], 1.0f / (-p.psi_s[i]), e1 - 1.0f, 1.0f / p.theta_s[i]);
      base[(kPairRowF2 + i) * STRIDE] = make_float4(p.hksat[i], b, 0.0f, 0.0f);
    }
    set_rootr(rootr, h);
  }
};

struct PairState {
  float h2o[NH], smp[NH]; /* the lane's four layers, local order */
  float zwt, wa, rnf_sum; /* per-cell scalars, held by both lanes */
};

__device__ __forceinline__ float pair_xor(float v) { return __shfl_xor_sync(kFullMask, v, 1); }

/* value of layer i (0-based, 0..7) of a per-lane array on both lanes of the pair; safe
 * inside regions where other pairs of the warp have diverged (pair mask) */
__device__ __forceinline__ float pair_fetch(const float (&a)[NH], int i, unsigned pmask, int lane) {
  H9_ASSERT(i >= 0 && i < NL);
  const int jl = (i < NH) ? i : NL - 1 - i;
  float v = a[0];
#pragma unroll
  for (int k = 1; k < NH; ++k)
    if (jl == k) v = a[k];
  return __shfl_sync(pmask, v, (lane & ~1) | (i >= NH ? 1 : 0));
}

/* all eight layers of a per-lane array, in layer order, on both lanes */
__device__ __forceinline__ void pair_gather(const float (&a)[NH], int h, unsigned pmask, float (&full)[NL]) {
#pragma unroll
  for (int j = 0; j < NH; ++j) {
    const float other = __shfl_xor_sync(pmask, a[j], 1);
    full[j] = h ? other : a[j];          /* layer j (0-based) lives in lane 0 at j */
    full[NL - 1 - j] = h ? a[j] : other; /* layer 7-j lives in lane 1 at j */
  }
}

template <class C>
__device__ __forceinline__ float pair_specific_yield(const C& c, int i, float zwtmm) {
  const float s_y = c.ths(i) * (1.0f - fast_pow(fmaf(zwtmm, c.inv_npsi(i), 1.0f), c.me(i)));
  return fmaxf(s_y, 0.02f);
}

/* Control flow as in hydrology_step_fast (h9_physics_fast.cuh): block A is one basic block up
 * to the junction solve; MODE kStepAllDeep has a straight-line tail, kStepGeneral the
 * first-iteration-with-selects tail and the rare restart on the looping code. */
template <int MODE, class C>
__device__ __forceinline__ uint32_t hydrology_step_pair(const Geo& g, const GeoDyn& gd, const PairGeo& pg,
                                                        const C& c, const DayFast& d, PairState& s,
                                                        StepOut& o, const int h, const int lane H9_TICKS_PARAM) {
  constexpr float kLog2e = 1.4426950408889634f;
  const unsigned pmask = 3u << (lane & ~1);
  uint32_t fault = 0;
  const float dt = g.dt, rdt = g.rdt;
  const bool odd = (h != 0);
  H9_TICK_START();

  /* ------------------------------ block A ------------------------------ */
  /* :141-151 */
  float theta[NH];
#pragma unroll
  for (int j = 0; j < NH; ++j) theta[j] = s.h2o[j] * pg.rdzw[j];
  const float wl0 = (s.h2o[0] + s.h2o[1]) + (s.h2o[2] + s.h2o[3]);
  const float pth0 = pair_xor(theta[0]);
  const float pth3 = pair_xor(theta[3]); /* the layer across the junction */
  const float w0 = (d.rain_dt + s.wa) + (wl0 + pair_xor(wl0));
  const float th_top = odd ? pth0 : theta[0]; /* theta(1) */
  const float4 mi = c.misc();                 /* Fmax, min hksat(1:3), rootr_col(1), 1/theta_s(1) */

  /* SurfaceRunoff :182-212 */
  const float fsat = mi.x * MathFast::ex2((-0.5f * kFff * kLog2e) * s.zwt);
  float qflx_surf = fsat * d.forc_rain;

  /* beta from the previous sub-step's smp :269-276 */
  const float4 rr4 = c.rootr4();
  const float rr[NH] = {rr4.x, rr4.y, rr4.z, rr4.w};
  float bl;
  {
    float bw[NH];
#pragma unroll
    for (int j = 0; j < NH; ++j) bw[j] = __saturatef(fmaf(s.smp[j], 1.0f / 150000.0f, pg.kbeta[j]));
    bl = fmaf(rr[0], bw[0], fmaf(rr[1], bw[1], fmaf(rr[2], bw[2], rr[3] * bw[3])));
  }
  const float beta = bl + pair_xor(bl);

  /* rsc :283-295, rss :325-331 */
  float rsc = (d.canopy_on && beta > 0.0f) ? d.rsc_num * MathFast::rcp(d.rsc_den0 * beta) : 1.0E6f;
  rsc = fmaxf(rsc, d.rsc_floor);
  const float rss_dry = d.litter10 * MathFast::ex2((35.63f * kLog2e) * (0.15f - th_top));
  const float rss_wet = fmaf(d.litter1000, 1.0f - th_top * mi.w, 10.0f);
  const float rss = (th_top <= 0.15f) ? rss_dry : rss_wet;

  /* two-source Penman-Monteith :344-389 (as in h9_physics_fast.cuh) */
  const float dPMc = fmaf(d.gamma, fmaf(rsc, d.inv_raa_rac, 1.0f), d.desatdT);
  const float dPMs = fmaf(d.gamma, fmaf(rss, d.inv_raa_ras, 1.0f), d.desatdT);
  const float Rs = fmaf(d.gamma, rss, d.dg_ras);
  const float Rc = fmaf(d.gamma, rsc, d.dg_rac);
  const float RsRc = Rs * Rc, RsRa = Rs * d.Ra, RcRa = Rc * d.Ra;
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
  const float evap_max1 = fmaxf(0.0f, fmaf(-qflx_tran_veg_col, mi.z, g.dz[1] * (th_top - kWatmin) * rdt));
  qflx_evap_grnd = fminf(evap_max1, qflx_evap_grnd);

  /* Infiltration :426-478 */
  const float qflx_in_soil = (d.forc_rain - qflx_surf) - qflx_evap_grnd;
  const float qinmax = (1.0f - fsat) * mi.y;
  const float qflx_infl_excess = fmaxf(0.0f, qflx_in_soil - qinmax);
  const float qflx_infl = qflx_in_soil - qflx_infl_excess;
  qflx_surf += qflx_infl_excess;

  /* SoilWater :492-508 */
  float zwtmm = 1000.0f * s.zwt;
  const bool deep = !(s.zwt <= g.zim[NL]);
  const float dz9 = deep ? (zwtmm - g.zc[NL]) : g.dz[NL];
  const float rous_early = pair_specific_yield(c, NL - 1, zwtmm);
  const float rrous_early = MathFast::rcp(rous_early);

  /* node arrays with the outer node at index 0 and the node across the junction at NH+1:
   * X[1+j] is local layer j */
  float zq[NH + 2], smpn[NH + 2], dsm[NH + 2], snode[NH];
  float tsum[NH + 1]; /* theta(k)+theta(k+1) of interface k = -1..3 */
  tsum[0] = theta[0] + theta[0];
#pragma unroll
  for (int k = 0; k < NH - 1; ++k) tsum[k + 1] = theta[k] + theta[k + 1];
  tsum[NH] = theta[NH - 1] + pth3;

  float rden_m1;
  { /* outer node of the odd lane: the aquifer layer :576-590 and node :645-650 (the even
     * lane runs the same instructions on its top layer's constants; the result is unused) */
    const float4 A = c.g0(0), B = c.g1(0);
    const float u = (zwtmm - g.zi[NL]) * A.y;
    const float a2 = 0.5f * (A.z - 1.0f), a3 = (1.0f / 3.0f) * (A.z - 2.0f), a4 = 0.25f * (A.z - 3.0f);
    const float r_series = A.z * fmaf(u * a2, fmaf(u * a3, fmaf(u, a4, 1.0f), 1.0f), 1.0f);
    const float r_direct = (fast_pow(1.0f + u, A.z) - 1.0f) * MathFast::rcp(u);
    const float r = (u < 0.03f) ? r_series : r_direct;
    const float coefA = A.w * (pg.zhi[0] - pg.zlo[0]);
    float vol = -coefA * A.y * r;
    vol = fminf(B.x, fmaxf(vol, 0.0f));
    const float z9 = fmaxf(kSmpmin, A.x * fast_pow(fmaxf(vol * B.y, 0.01f), -B.z));
    zq[0] = deep ? z9 : 0.0f;
    const float s_node = fminf(1.0f, fmaxf(0.5f * fmaf(theta[0], B.y, 1.0f), 0.01f));
    smpn[0] = fmaxf(kSmpmin, A.x * fast_pow(s_node, -B.z));
    const float zc9 = 0.5f * (zwtmm + g.zc[NL]);
    const float den9 = zc9 - g.zc[NL];
    const float r9 = MathFast::rcp(s_node * den9);
    dsm[0] = B.w * smpn[0] * (r9 * den9);
    rden_m1 = r9 * s_node; /* 1/(zc(9)-zc(8)) of interface -1 */
  }

#pragma unroll
  for (int j = 0; j < NH; ++j) { /* equilibrium profile :517-573 and matric potential :626-639 */
    const float4 A = c.g0(j), B = c.g1(j);
    const float t0 = fast_pow(fmaxf(fmaf(zwtmm - pg.zlo[j], A.y, 1.0f), 1.0f), A.z);
    const float ti = fast_pow(fmaxf(fmaf(zwtmm - pg.zhi[j], A.y, 1.0f), 1.0f), A.z);
    float vol = fmaf(A.w, ti - t0, B.x * fmaxf(fmaf(-zwtmm, pg.rdzl[j], pg.zhr[j]), 0.0f));
    vol = fminf(B.x, fmaxf(vol, 0.0f));
    zq[1 + j] = fmaxf(kSmpmin, A.x * fast_pow(fmaxf(vol * B.y, 0.01f), -B.z));
    const float s_node = fminf(1.0f, fmaxf(theta[j] * B.y, 0.01f));
    const float sm = fmaxf(kSmpmin, A.x * fast_pow(s_node, -B.z));
    s.smp[j] = sm;
    smpn[1 + j] = sm;
    snode[j] = s_node;
    dsm[1 + j] = B.w * sm; /* (-b/ths)*smp, still to be divided by s_node */
  }
#pragma unroll
  for (int k = 0; k < NH; k += 2) { /* two reciprocals from one MUFU; s_node is in [0.01, 1] */
    const float r = MathFast::rcp(snode[k] * snode[k + 1]);
    dsm[1 + k] *= r * snode[k + 1];
    dsm[2 + k] *= r * snode[k];
  }
  /* the node across the junction */
  smpn[NH + 1] = pair_xor(smpn[NH]);
  zq[NH + 1] = pair_xor(zq[NH]);
  dsm[NH + 1] = pair_xor(dsm[NH]);

  /* interfaces k = -1..3: conductivity :598-621, inward flux and its two derivatives :661-753 */
  float F[NH + 1], Fa[NH + 1], Fb[NH + 1];
#pragma unroll
  for (int k = -1; k < NH; ++k) {
    const float4 G = c.gi(k);
    const float s1 = fminf(1.0f, tsum[k + 1] * G.x);
    const float s2 = G.y * fast_pow(s1, G.z);
    const float hk = s1 * s2;
    const float dhk = G.w * s2;
    const float rden = (k < 0) ? rden_m1 : pg.rden[k < 0 ? 0 : k];
    const float num = (smpn[k + 2] - smpn[k + 1]) - (zq[k + 2] - zq[k + 1]);
    const float nd = num * dhk;
    F[k + 1] = -hk * num * rden;
    Fa[k + 1] = fmaf(hk, dsm[k + 1], -nd) * rden;  /* dF/dtheta(k)   */
    Fb[k + 1] = -fmaf(hk, dsm[k + 2], nd) * rden;  /* dF/dtheta(k+1) */
  }
  /* outer boundary: infiltration for the even lane (:661-672), the aquifer interface for the
   * odd lane (:737-753), inert when the water table is inside the column (:732-735) */
  const bool aq_on = odd && deep;
  const float Fm1 = aq_on ? F[0] : (odd ? 0.0f : qflx_infl);
  const float Fam1 = aq_on ? Fa[0] : 0.0f;
  const float Fbm1 = aq_on ? Fb[0] : 0.0f;

  /* elimination from the outer node towards the junction; row j:
   *   -Fa(j-1) dw(j-1) + (dz/dt + Fa(j) - Fb(j-1)) dw(j) + Fb(j) dw(j+1) = F(j-1) - F(j) - sink(j) */
  float u[NH + 1], gam[NH + 1]; /* index j+1 */
  float minpiv;
  float rbet;
  {
    const float bet = odd ? fmaf(dz9, rdt, Fam1) : 1.0f; /* aquifer row 9 :755-799 */
    minpiv = fabsf(bet);
    rbet = MathFast::rcp(bet);
    u[0] = odd ? -Fm1 * rbet : 0.0f;
    gam[0] = Fbm1 * rbet;
  }
  {
    const float rmx = fmaf(-qflx_tran_veg_col, rr[0], Fm1 - F[1]);
    const float dg = (pg.dzdt[0] - Fbm1) + Fa[1];
    const float bet = fmaf(Fam1, gam[0], dg);
    fault |= (!odd && bet == 0.0f) ? FAULT_PIVOT1 : 0u; /* bmx(1) == 0 :806 */
    minpiv = fminf(minpiv, fabsf(bet));
    rbet = MathFast::rcp(bet);
    u[1] = fmaf(Fam1, u[0], rmx) * rbet;
    gam[1] = Fb[1] * rbet;
  }
#pragma unroll
  for (int j = 1; j < NH; ++j) {
    const float rmx = fmaf(-qflx_tran_veg_col, rr[j], F[j] - F[j + 1]);
    const float dg = (pg.dzdt[j] - Fb[j]) + Fa[j + 1];
    const float bet = fmaf(Fa[j], gam[j], dg);
    minpiv = fminf(minpiv, fabsf(bet));
    rbet = MathFast::rcp(bet);
    u[j + 1] = fmaf(Fa[j], u[j], rmx) * rbet;
    gam[j + 1] = Fb[j + 1] * rbet;
  }
  /* junction: dw(3) + gam dw'(3) = u(3) on both sides */
  float dw[NH + 1]; /* index j+1 */
  {
    const float pu = pair_xor(u[NH]);
    const float pgm = pair_xor(gam[NH]);
    const float den = fmaf(-gam[NH], pgm, 1.0f);
    minpiv = fminf(minpiv, fabsf(den));
    dw[NH] = fmaf(-gam[NH], pu, u[NH]) * MathFast::rcp(den);
  }
#pragma unroll
  for (int j = NH - 1; j >= 0; --j) dw[j] = fmaf(-gam[j], dw[j + 1], u[j]);
  const float dwat9 = __shfl_sync(kFullMask, dw[0], lane | 1); /* dwat2(9), held by the odd lane */
  fault |= (minpiv == 0.0f) ? FAULT_PIVOT2 : 0u;

  /* inputs of the tail of a cell whose table is inside the column that do not depend on the
   * solve (see hydrology_step_fast): jwt as a count, the Darcy recharge :866-895, the first
   * specific yield.  The warp is converged here: full-mask shuffles. */
  int jc = NL - 1;
  float qcharge_early = 0.0f, sy_first = 0.02f, rsy_first = 50.0f;
  if (MODE == kStepGeneral) {
    /* (as an add tree: the count sits on the dependent chain of the tail) */
    const int cnt = (((s.zwt > g.zim[1]) + (s.zwt > g.zim[2])) + ((s.zwt > g.zim[3]) + (s.zwt > g.zim[4]))) +
                    (((s.zwt > g.zim[5]) + (s.zwt > g.zim[6])) + ((s.zwt > g.zim[7]) + (s.zwt > g.zim[8])));
    jc = (cnt < NL) ? cnt : NL - 1;
    const int jm = (jc > 1 ? jc : 1) - 1;
    float smp_own[NH], zq_own[NH];
#pragma unroll
    for (int j = 0; j < NH; ++j) {
      smp_own[j] = s.smp[j];
      zq_own[j] = zq[1 + j];
    }
    const float th_j = pair_fetch(theta, jc, kFullMask, lane);
    const float smp1 = fmaxf(kSmpmin, pair_fetch(smp_own, jm, kFullMask, lane));
    const float zq_j = pair_fetch(zq_own, jm, kFullMask, lane);
    const float s1 = fminf(1.0f, fmaxf(th_j * c.inv_ths(jc), 0.01f));
    const float ka = c.hks(jc) * fast_pow(s1, fmaf(2.0f, c.bsw(jc), 3.0f));
    const float wh = smp1 - zq_j;
    const float denom = (jc == 0) ? (zwtmm + 1.0f) : (zwtmm - gd.zc(jc)) * 2.0f;
    qcharge_early = ka * wh * MathFast::rcp(denom);
    qcharge_early = fminf(g.q10_hi, fmaxf(g.q10_lo, qcharge_early));
    sy_first = pair_specific_yield(c, jc, zwtmm);
    rsy_first = MathFast::rcp(sy_first);
  }

  H9_TICK(MODE == kStepGeneral ? 5 : 0); /* block A: 0 = all-deep step, 5 = general step */
  /* ------------------------------ block B ------------------------------ */
  /* state update :845-850 */
#pragma unroll
  for (int j = 0; j < NH; ++j) s.h2o[j] = fmaf(dw[j + 1], pg.dz[j], s.h2o[j]);
  /* layer 8 (the odd lane's j = 0) as the even lane will need it for the trigger at the end of
   * the tail, sent now: what the aquifer spills into it later both lanes compute themselves */
  const float h8_sent = pair_xor(s.h2o[0]);
  float spill8 = 0.0f;

  float rsub_top;
  int jwt = NL;
  float zwt_j;
  float imb = 0.0f, rnf_inc = 0.0f;
  const float rnf_sum_in = s.rnf_sum;
  uint32_t fault_bal = 0, fault_tail = 0;
  const float4 capv = c.cap();
  const float cp[NH] = {capv.x, capv.y, capv.z, capv.w};
  /* balance :1221-1244 and the runoff sum :1282-1283; the column sum crosses the pair */
  auto balance = [&](float qflx_rsub_sat, unsigned mask) {
    const float wl1 = (s.h2o[0] + s.h2o[1]) + (s.h2o[2] + s.h2o[3]);
    const float w1 =
        fmaf((qflx_surf + qflx_evap_grnd + qflx_tran_veg_col) + rsub_top + qflx_rsub_sat, dt, s.wa) +
        (wl1 + __shfl_xor_sync(mask, wl1, 1));
    imb = w1 - w0;
    fault_bal = (!(fabsf(imb) <= 0.1f)) ? FAULT_IMBAL : 0u;
    const float r1 = qflx_surf * dt, r2 = rsub_top * dt;
    rnf_inc = r1 + r2;
    s.rnf_sum = fmaf(rsub_top, dt, fmaf(qflx_surf, dt, rnf_sum_in));
  };
  auto excess_local = [&]() { /* max over the lane's layers of (water - capacity), :1131-1148 */
    return fmaxf(fmaxf(s.h2o[0] - cp[0], s.h2o[1] - cp[1]), fmaxf(s.h2o[2] - cp[2], s.h2o[3] - cp[3]));
  };
  auto lowest_local = [&]() {
    return fminf(fminf(s.h2o[0], s.h2o[1]), fminf(s.h2o[2], s.h2o[3]));
  };
  /* the dryness repair :1161-1174 for a table below the column (it cannot move zwt, :1166),
   * straight-line over the pair: layers 1..4 top-down in the even lane, the deficit of layer 4
   * handed across the junction, layers 5..7 in the odd lane (its local order is bottom-up);
   * `on` masks it per cell.  With D_j = watmin - h2o_j the reference's loop is
   * xs_j = max(0, D_j + xs_{j-1}), h2o_j = (h2o_j - xs_{j-1}) + xs_j: two dependent operations
   * per layer in the even lane, and in the odd lane every xs as max(B, A + p) of the deficit p
   * that arrives over the junction (A, B from the lane's own D, ready before the shuffle
   * lands), so three operations follow the shuffle instead of nine.  A pass that finds no layer
   * below the floor (the usual case) leaves every h2o bit for bit. */
  /* Returns whether layer 8 ends below the floor (:1181-1198, the looping repair of the branch):
   * the odd lane sends the three numbers that decide it while the even lane sends its deficit,
   * and both lanes evaluate the same expression on the same operands, so the pair agrees on
   * the branch without a second shuffle behind the pass. */
  auto dryness_pass = [&](bool on) -> bool {
    float D[NH];
#pragma unroll
    for (int j = 0; j < NH; ++j) D[j] = kWatmin - s.h2o[j];
    /* even lane: layers 1..4 = local 0..3 */
    const float x0 = fmaxf(D[0], 0.0f);
    const float x1 = fmaxf(D[1] + x0, 0.0f);
    const float x2 = fmaxf(D[2] + x1, 0.0f);
    const float x3 = fmaxf(D[3] + x2, 0.0f);
    const float xe = (on && !odd) ? x3 : 0.0f;
    /* odd lane: layers 5, 6, 7 = local 3, 2, 1; layer 8 (local 0) only gives */
    const float a6 = D[3] + D[2], b6 = fmaxf(D[2], 0.0f);
    const float a7 = D[1] + a6, b7 = fmaxf(D[1] + b6, 0.0f);
    const float r_b = pair_xor(b7);
    const float r_x = pair_xor(odd ? a7 : xe); /* even lane: the odd lane's a7; odd lane: layer 4's deficit */
    const float p = odd ? r_x : xe;
    const float y7 = fmaxf(odd ? b7 : r_b, (odd ? a7 : r_x) + p);
    const bool dry8 = on && ((odd ? s.h2o[0] : h8_sent + spill8) - y7) < kWatmin;
    const float y5 = fmaxf(D[3] + p, 0.0f);
    const float y6 = fmaxf(b6, a6 + p);
    const float add[NH] = {odd ? 0.0f : x0, odd ? y7 : x1, odd ? y6 : x2, odd ? y5 : x3};
    const float sub[NH] = {odd ? y7 : 0.0f, odd ? y6 : x0, odd ? y5 : x1, odd ? p : x2};
#pragma unroll
    for (int j = 0; j < NH; ++j) s.h2o[j] = (s.h2o[j] - (on ? sub[j] : 0.0f)) + (on ? add[j] : 0.0f);
    return dry8;
  };
  /* excess cascade :1131-1152 and dryness repair :1161-1205 as the reference orders them, on
   * the gathered column; both lanes of the pair take the branch together */
  auto repair = [&]() {
    float qflx_rsub_sat = 0.0f, xs = 0.0f;
    float w[NL];
    pair_gather(s.h2o, h, pmask, w);
#pragma unroll
    for (int i = NL - 1; i >= 1; --i) {
      const float cap = fmaxf(0.01f, c.ths(i)) * g.dz[i + 1];
      const float xsi = fmaxf(w[i] - cap, 0.0f);
      w[i] = fminf(cap, w[i]);
      w[i - 1] += xsi;
    }
    const float cap1 = fmaxf(0.0f, c.ths(0) * g.dz[1]);
    const float xs1 = fmaxf(fmaxf(w[0], 0.0f) - cap1, 0.0f);
    w[0] = fminf(cap1, w[0]);
    qflx_rsub_sat = xs1 * rdt;
    if (jwt < 0) jwt = find_jwt(g, zwt_j);
#pragma unroll
    for (int i = 0; i < NL - 1; ++i) {
      xs = fmaxf(kWatmin - w[i], 0.0f);
      if (xs > 0.0f && i + 1 == jwt) s.zwt += xs * MathFast::rcp(fmaxf(0.01f, c.ths(i))) * 0.001f;
      w[i] += xs;
      w[i + 1] -= xs;
    }
    xs = 0.0f;
    if (w[NL - 1] < kWatmin) { /* :1181-1198 */
      xs = kWatmin - w[NL - 1];
      bool done = false;
#pragma unroll
      for (int j = NL - 2; j >= 0; --j) {
        if (!done) {
          const float avail = fmaxf(w[j] - kWatmin - xs, 0.0f);
          const float take = (avail >= xs) ? xs : avail;
          done = (avail >= xs);
          w[NL - 1] += take;
          w[j] -= take;
          xs = done ? 0.0f : xs - take;
        }
      }
    }
    w[NL - 1] += xs; /* :1205 */
#pragma unroll
    for (int j = 0; j < NH; ++j) s.h2o[j] = odd ? w[NL - 1 - j] : w[j];
    rsub_top -= xs * rdt; /* :1211 */
    balance(qflx_rsub_sat, pmask);
  };
  /* Recharge :896-904, Drainage :946-951, baseflow :1048-1058 for a table below the column */
  auto deep_tail = [&]() {
    const float qcharge = dwat9 * dz9 * rdt;
    s.wa = fmaf(qcharge, dt, s.wa);
    s.zwt = fmaf(-(qcharge * dt * 0.001f), rrous_early, s.zwt);
    zwtmm = 1000.0f * s.zwt; /* :1015 */
    rsub_top = 5.5E-3f * MathFast::ex2((-kFff * kLog2e) * s.zwt);
    const float rous = pair_specific_yield(c, NL - 1, zwtmm);
    s.wa = fmaf(-rsub_top, dt, s.wa);
    s.zwt = fmaf(rsub_top * dt * 0.001f, MathFast::rcp(rous), s.zwt);
    spill8 = fmaxf(0.0f, s.wa - 5000.0f);
    s.h2o[0] += odd ? spill8 : 0.0f; /* layer 8 is the odd lane's j = 0 */
    s.wa = fminf(s.wa, 5000.0f);
    zwt_j = s.zwt;
  };
  /* the data-dependent tail of a table inside the column, as the reference loops it */
  auto shallow_tail = [&]() {
    const int jfirst = jwt + 1;
    float qcharge_tot = qcharge_early * dt;
    if (qcharge_tot > 0.0f) {
      for (int I = jwt + 1; I >= 1; --I) {
        const float s_y = (I == jfirst) ? sy_first : pair_specific_yield(c, I - 1, zwtmm);
        const float ql = fmaxf(fminf(qcharge_tot, s_y * (zwtmm - gd.zi(I - 1))), 0.0f);
        s.zwt -= ql * MathFast::rcp(s_y) * 0.001f;
        qcharge_tot -= ql;
        if (qcharge_tot <= 0.0f) break;
      }
    } else {
      for (int I = jwt + 1; I <= NL; ++I) {
        const float s_y = (I == jfirst) ? sy_first : pair_specific_yield(c, I - 1, zwtmm);
        const float ql = fminf(fmaxf(qcharge_tot, -s_y * (gd.zi(I) - zwtmm)), 0.0f);
        qcharge_tot -= ql;
        if (qcharge_tot >= 0.0f) {
          s.zwt -= ql * MathFast::rcp(s_y) * 0.001f;
          break;
        }
        s.zwt = gd.zim(I);
      }
      if (qcharge_tot > 0.0f) s.zwt -= qcharge_tot * 0.001f * rrous_early;
    }
    jwt = find_jwt(g, s.zwt);
    zwtmm = 1000.0f * s.zwt; /* :1015 */
    rsub_top = 5.5E-3f * MathFast::ex2((-kFff * kLog2e) * s.zwt);
    const float rous = pair_specific_yield(c, NL - 1, zwtmm);
    if (jwt == NL) {
      s.wa = fmaf(-rsub_top, dt, s.wa);
      s.zwt = fmaf(rsub_top * dt * 0.001f, MathFast::rcp(rous), s.zwt);
      s.h2o[0] += odd ? fmaxf(0.0f, s.wa - 5000.0f) : 0.0f;
      s.wa = fminf(s.wa, 5000.0f);
    } else {
      float rsub_top_tot = -rsub_top * dt;
      if (rsub_top_tot > 0.0f) {
        fault_tail |= FAULT_RSUB;
      } else {
        for (int I = jwt + 1; I <= NL; ++I) {
          const float s_y = pair_specific_yield(c, I - 1, zwtmm);
          const float rl = fminf(fmaxf(rsub_top_tot, -(s_y * (gd.zi(I) - zwtmm))), 0.0f);
          const int jl = (I - 1 < NH) ? I - 1 : NL - I; /* local index of layer I */
          const bool mine = ((I - 1 >= NH) == odd);
#pragma unroll
          for (int k = 0; k < NH; ++k) s.h2o[k] = (mine && k == jl) ? s.h2o[k] + rl : s.h2o[k];
          rsub_top_tot -= rl;
          if (rsub_top_tot >= 0.0f) {
            s.zwt -= rl * MathFast::rcp(s_y) * 0.001f;
            break;
          }
          s.zwt = gd.zim(I);
        }
        s.zwt -= rsub_top_tot * 0.001f * MathFast::rcp(rous); /* :1100-1101, unconditional (G11) */
        s.wa += rsub_top_tot;
      }
      jwt = -1;
    }
    zwt_j = s.zwt;
  };
  /* merge the pair's trigger and fault bits: one shuffle */
  auto merge_flags = [&](bool fix_local, unsigned mask) {
    uint32_t flags = fault | fault_tail | (fix_local ? 0x100u : 0u);
    flags |= __shfl_xor_sync(mask, flags, 1);
    fault = flags & 0xFFu;
    return (flags & 0x100u) != 0u;
  };

  if (MODE == kStepAllDeep) {
    deep_tail();
    s.zwt = fminf(80.0f, fmaxf(0.0f, zwt_j)); /* :1122-1123 */
    const float over = excess_local();
    float h_pre[NH];
#pragma unroll
    for (int j = 0; j < NH; ++j) h_pre[j] = s.h2o[j];
    /* the balance before the pass: the pass moves water between layers of the column, its sum
     * is the same up to rounding (bit for bit when the pass finds nothing to do), and the pass
     * and the balance are the two long chains at the end of the sub-step */
    const bool fix_early = merge_flags(over > 0.0f, kFullMask);
    balance(0.0f, kFullMask);
    /* layer 8 (the odd lane's j = 0) is not reached by the pass: :1181-1198 is in the branch */
    if (dryness_pass(true) || fix_early) {
#pragma unroll
      for (int j = 0; j < NH; ++j) s.h2o[j] = h_pre[j]; /* the cascade comes first */
      repair();
    }
  } else {
    /* general straight-line tail, see hydrology_step_fast */
    const float zwt0 = s.zwt, wa0 = s.wa;
    float h_bs[NH];
#pragma unroll
    for (int j = 0; j < NH; ++j) h_bs[j] = s.h2o[j];
    const float qcharge = deep ? dwat9 * dz9 * rdt : qcharge_early;
    const float qtot = qcharge * dt;
    H9_ASSERT(jc >= 0 && jc + 1 <= NL);
    const bool up = qtot > 0.0f;
    const float ql_up = fmaxf(fminf(qtot, sy_first * (zwtmm - gd.zi(jc))), 0.0f);
    const float ql_dn = fminf(fmaxf(qtot, -sy_first * (gd.zi(jc + 1) - zwtmm)), 0.0f);
    const float ql = up ? ql_up : ql_dn;
    const float qrem = qtot - ql;
    const float zmove = zwt0 - ql * rsy_first * 0.001f;
    const bool done_dn = qrem >= 0.0f;
    const float zwt1_sh = (up || done_dn) ? zmove : gd.zim(jc + 1);
    const bool more1 = up ? (qrem > 0.0f && jc >= 1) : (!done_dn && jc + 2 <= NL);
    const float zwt1 = deep ? fmaf(-(qcharge * dt * 0.001f), rrous_early, zwt0) : zwt1_sh;
    const float wa1 = deep ? fmaf(qcharge, dt, wa0) : wa0;
    const float zwtmm1 = 1000.0f * zwt1; /* :1015 */
    rsub_top = 5.5E-3f * MathFast::ex2((-kFff * kLog2e) * zwt1);
    const float rous = pair_specific_yield(c, NL - 1, zwtmm1);
    const float rrous = MathFast::rcp(rous);
    const int cnt2 = (((zwt1 > g.zim[1]) + (zwt1 > g.zim[2])) + ((zwt1 > g.zim[3]) + (zwt1 > g.zim[4]))) +
                     (((zwt1 > g.zim[5]) + (zwt1 > g.zim[6])) + ((zwt1 > g.zim[7]) + (zwt1 > g.zim[8])));
    const bool isA = deep || cnt2 == NL;
    const int j2c = (cnt2 < NL) ? cnt2 : NL - 1;
    H9_ASSERT(j2c >= 0 && j2c + 1 <= NL);
    const float sy2 = pair_specific_yield(c, j2c, zwtmm1);
    const float rtot = -rsub_top * dt;
    const float rl = fminf(fmaxf(rtot, -(sy2 * (gd.zi(j2c + 1) - zwtmm1))), 0.0f);
    const float rrem = rtot - rl;
    const bool done2 = rrem >= 0.0f;
    const float zwt_b = done2 ? zwt1 - rl * MathFast::rcp(sy2) * 0.001f : gd.zim(j2c + 1);
    const bool more2 = !isA && ((!done2 && j2c + 2 <= NL) || rtot > 0.0f);
    const float zwt_c = zwt_b - rrem * 0.001f * rrous;
    const float wa_c = wa1 + rrem;
    const float wa_a = fmaf(-rsub_top, dt, wa1);
    const float zwt_a = fmaf(rsub_top * dt * 0.001f, rrous, zwt1);
    zwt_j = isA ? zwt_a : zwt_c;
    jwt = isA ? NL : -1;
    const float wa2 = isA ? wa_a : wa_c;
    {
      const int jl = (j2c < NH) ? j2c : NL - 1 - j2c; /* local index of layer j2c+1 */
      const bool mine = !isA && ((j2c >= NH) == odd);
#pragma unroll
      for (int k = 0; k < NH; ++k) s.h2o[k] = (mine && k == jl) ? s.h2o[k] + rl : s.h2o[k];
    }
    spill8 = isA ? fmaxf(0.0f, wa2 - 5000.0f) : 0.0f;
    s.h2o[0] += odd ? spill8 : 0.0f;
    s.wa = isA ? fminf(wa2, 5000.0f) : wa2;
    /* clamp, triggers, the dryness pass where it cannot move the table, balance (see the general
     * tail of hydrology_step_fast) */
    s.zwt = fminf(80.0f, fmaxf(0.0f, zwt_j));
    const float over = excess_local();
    const float lowest = lowest_local();
    const bool fix_local = (more1 && !deep) || more2 || over > 0.0f || (!isA && lowest < kWatmin);
    const bool fix_early = merge_flags(fix_local, kFullMask);
    balance(0.0f, kFullMask); /* before the pass: see the all-deep tail */
    if (dryness_pass(isA) || fix_early) { /* rare: the looping code from the saved state */
      s.zwt = zwt0;
      s.wa = wa0;
#pragma unroll
      for (int j = 0; j < NH; ++j) s.h2o[j] = h_bs[j];
      zwtmm = 1000.0f * zwt0;
      if (deep) {
        deep_tail();
      } else {
        jwt = jc;
        shallow_tail();
      }
      s.zwt = fminf(80.0f, fmaxf(0.0f, zwt_j));
      fault |= fault_tail;
      const bool fix2 = merge_flags(excess_local() > 0.0f || lowest_local() < kWatmin, pmask);
      balance(0.0f, pmask);
      if (fix2) repair();
    }
  }

  H9_TICK(MODE == kStepGeneral ? 2 : 1); /* tail: 1 = all-deep step, 2 = general step */
#ifdef H9_CYCLE_BUDGET
  tk.acc[MODE == kStepGeneral ? 4 : 3] += 1u; /* sub-steps on each variant */
#endif
  o.qflx_tran_veg_col = qflx_tran_veg_col;
  o.qflx_evap_grnd = qflx_evap_grnd;
  o.rnf_inc = rnf_inc;
  o.imbalance = imb;
  o.jwt = (jwt < 0) ? find_jwt(g, zwt_j) : jwt;
  return fault | fault_bal;
}

} /* namespace h9 */
#endif
