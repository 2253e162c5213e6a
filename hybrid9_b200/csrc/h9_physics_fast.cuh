/*
 * h9_physics_fast.cuh -- H9_MATH_FAST implementation of one HYDROLOGY sub-step
 * (HYDROLOGY.f90:141-1283), restructured for the B200's issue and MUFU limits.
 *
 * Same algorithm and the same branch semantics as hydrology_step<M> in
 * h9_physics.h (which follows the Fortran operation by operation); what
 * changes is the arithmetic around it:
 *   - every per-cell, per-layer quantity that does not depend on the state
 *     (1/(-psi_s), 1-1/bsw, psi_s*theta_s/(1-1/bsw)/dz, 1/theta_s, 1/(theta_s(I)+
 *     theta_s(I+1)), MINVAL(hksat(1:3)) ...) is computed once per launch and kept
 *     in shared memory, one column per thread (no bank conflicts, no barriers);
 *   - divisions by geometry or per-day constants are multiplications by
 *     reciprocals prepared on the host / once per day;
 *   - pow(a,b) = ex2(b*lg2(a)), exp(x) = ex2(x*log2 e), 1/x = rcp.approx: 2 MUFU
 *     per pow; the two Cc/Cs fractions of HYDROLOGY.f90:362-366 share one
 *     denominator; interface fluxes are computed once per interface (the
 *     reference writes each twice, as qout(I) and qin(I+1));
 *   - the data-dependent Drainage loops are real loops over the shared-memory
 *     constants instead of 8-fold predicated unrolls (code size, I-cache);
 *   - latency: layers are processed in pairs, each pair followed at once by the interface
 *     fluxes and the Thomas rows it completes (the sweep's reciprocal chain runs behind the
 *     powers of the deeper layers), and the parts of Recharge/Drainage that do not depend on
 *     the solve are issued before it.
 * Results differ from the exact mode at rounding level only; the deviation is
 * measured against the oracle and the FP32 noise floor in tests/test_gpu_parity.py.
 *
 * This header holds the table, the per-day constants and the sub-step as the kernels with at
 * most two warps per scheduler run it (small shards of a multi-GPU split): control flow laid
 * out for a lone warp, see hydrology_step_fast below.  The sub-step of the 128-register
 * throughput build is in h9_physics_fast_tp.cuh (round 1's control flow, same arithmetic, same
 * bits); the two-lanes-per-cell sub-step in h9_physics_pair.cuh.
 */
#ifndef H9_PHYSICS_FAST_CUH
#define H9_PHYSICS_FAST_CUH

#include "h9_physics.h"

namespace h9 {

/* Cycle budget of one sub-step (profiles/r02/cycle_budget_*.txt): built only with
 * -DH9_CYCLE_BUDGET (tools/cycle_budget.py), never in the product library.  H9_TICK(k) closes
 * segment k: it reads %clock (a scheduling fence: segments no longer overlap, so their sum is
 * an upper bound of the unfenced sub-step) and adds the cycles since the previous tick. */
#ifdef H9_CYCLE_BUDGET
constexpr int kTickSegs = 12;
struct Ticks {
  unsigned acc[kTickSegs];
  unsigned last;
  __device__ __forceinline__ void start() {
    asm volatile("mov.u32 %0, %%clock;" : "=r"(last)::"memory");
  }
  __device__ __forceinline__ void tick(int k) {
    unsigned t;
    asm volatile("mov.u32 %0, %%clock;" : "=r"(t)::"memory");
    acc[k] += t - last;
    last = t;
  }
};
#define H9_TICKS_PARAM , Ticks& tk
#define H9_TICKS_ARG , tk
#define H9_TICK(k) tk.tick(k)
#define H9_TICK_START() tk.start()
#if H9_CYCLE_BUDGET >= 2 /* coarse: block A is left unfenced inside */
#define H9_TICK_FINE(k)
#else
#define H9_TICK_FINE(k) tk.tick(k)
#endif
#else
#define H9_TICKS_PARAM
#define H9_TICKS_ARG
#define H9_TICK(k)
#define H9_TICK_FINE(k)
#define H9_TICK_START()
#endif

/* The per-cell constant table: 27 float4 per cell.  Three groups per soil layer, so that one
 * 128-bit shared-memory load brings the four constants a section of the sub-step needs:
 *   G0 = (psi_s, 1/(-psi_s), 1-1/bsw, psi_s*theta_s/(1-1/bsw)/dz)      equilibrium profile
 *   G1 = (theta_s, 1/theta_s, bsw, hksat)                              both sections
 *   G2 = (its = 1/(theta_s(I)+theta_s(I+1)), psi_s, (2*bsw+3)*its, -bsw/theta_s)   hk section
 * then S = (Fmax, MINVAL(hksat(1:3)), -, -) and rootr_col(1:4), rootr_col(5:8) (beta and the
 * sinks read all eight: two loads). */
constexpr int kFastGroups = 3;
constexpr int kFastFloatsPerCell = (kFastGroups * NL + 3) * 4;

/* one thread's column: group gi of layer i lives at base[(gi*NL + i) * STRIDE] (float4 units);
 * STRIDE = threads per block when the table is in shared memory (consecutive threads ->
 * consecutive 16-byte slots: conflict-free 128-bit accesses), 1 for a local array */
template <int STRIDE>
struct CellTable {
  float4* base;
  __device__ __forceinline__ float4 g0(int i) const { return base[(0 * NL + i) * STRIDE]; }
  __device__ __forceinline__ float4 g1(int i) const { return base[(1 * NL + i) * STRIDE]; }
  __device__ __forceinline__ float4 g2(int i) const { return base[(2 * NL + i) * STRIDE]; }
  __device__ __forceinline__ float comp(int gi, int i, int k) const {
    H9_ASSERT(gi >= 0 && gi <= kFastGroups && i >= 0 && i < NL && k >= 0 && k < 4);
    return reinterpret_cast<const float*>(&base[(gi * NL + i) * STRIDE])[k];
  }
  __device__ __forceinline__ float psi(int i) const { return comp(0, i, 0); }
  __device__ __forceinline__ float inv_npsi(int i) const { return comp(0, i, 1); }
  __device__ __forceinline__ float e1(int i) const { return comp(0, i, 2); }
  __device__ __forceinline__ float coef3(int i) const { return comp(0, i, 3); }
  __device__ __forceinline__ float ths(int i) const { return comp(1, i, 0); }
  __device__ __forceinline__ float inv_ths(int i) const { return comp(1, i, 1); }
  __device__ __forceinline__ float bsw(int i) const { return comp(1, i, 2); }
  __device__ __forceinline__ float hks(int i) const { return comp(1, i, 3); }
  __device__ __forceinline__ float inv_ths_sum(int i) const { return comp(2, i, 0); }
  __device__ __forceinline__ float4 rootr4(int h) const { return base[(kFastGroups * NL + 1 + h) * STRIDE]; }
  __device__ __forceinline__ float rootr(int i) const { return comp(kFastGroups, 1 + (i >> 2), i & 3); }
  __device__ __forceinline__ void set_rootr(int i, float v) const {
    reinterpret_cast<float*>(&base[(kFastGroups * NL + 1 + (i >> 2)) * STRIDE])[i & 3] = v;
  }
  __device__ __forceinline__ float fmax() const { return comp(kFastGroups, 0, 0); }
  __device__ __forceinline__ float hksmin3() const { return comp(kFastGroups, 0, 1); }

  /* fill from the cell's parameters (once per launch) */
  __device__ __forceinline__ void init(const Geo& g, const Params& p, const float (&rootr)[NL]) const {
    const float hmin = fminf(fminf(p.hksat[0], p.hksat[1]), p.hksat[2]);
#pragma unroll
    for (int i = 0; i < NL; ++i) {
      const int ip = (i + 1 < NL) ? i + 1 : NL - 1;
      const float psi = p.psi_s[i], ths = p.theta_s[i], b = p.bsw[i];
      const float e1 = 1.0f - 1.0f / b;
      base[(0 * NL + i) * STRIDE] = make_float4(psi, 1.0f / (-psi), e1, psi * ths / e1 / (g.zi[i + 1] - g.zi[i]));
      base[(1 * NL + i) * STRIDE] = make_float4(ths, 1.0f / ths, b, p.hksat[i]);
      const float its = 1.0f / (ths + p.theta_s[ip]);
      base[(2 * NL + i) * STRIDE] = make_float4(its, psi, (2.0f * b + 3.0f) * its, -b / ths);
    }
    base[(kFastGroups * NL + 1) * STRIDE] = make_float4(rootr[0], rootr[1], rootr[2], rootr[3]);
    base[(kFastGroups * NL + 2) * STRIDE] = make_float4(rootr[4], rootr[5], rootr[6], rootr[7]);
    base[(kFastGroups * NL) * STRIDE] = make_float4(p.fmax, hmin, 0.0f, 0.0f);
  }
};

struct FastState {
  float h2o[NL], smp[NL];
  float zwt, wa, rnf_sum;
};

/* per-day constants of the fast path: Day plus the reciprocals the sub-step multiplies by */
struct DayFast {
  float forc_rain, rain_dt;
  float desatdT, gamma, dg, VDD, A;
  float rsc_num, rsc_den0 /* 2*LAI*p28 */, rsc_floor;
  float pmc_num, pms_num, inv_raa_rac, inv_raa_ras, inv_rac, inv_ras;
  float Ra, dg_ras, dg_rac;
  float raa_inv_rhocp, rhocp, lec_a, les_a, k_lamb /* 1e3/(rhow*lamb) */;
  float litter10, litter1000;
  bool canopy_on;
};

__device__ __forceinline__ void day_setup_fast(const Geo& g, const Forcing& f, float lai,
                                                float lai_litter, DayFast& o, float& tas) {
  Day d;
  day_setup<MathFast>(g, f, lai, lai_litter, d);
  tas = d.tas;
  o.forc_rain = d.forc_rain;
  o.rain_dt = d.rain_dt;
  o.desatdT = d.desatdT;
  o.gamma = d.gamma;
  o.dg = d.dg;
  o.VDD = d.VDD;
  o.A = d.A;
  o.rsc_num = d.rsc_num;
  o.rsc_den0 = d.lai2 * d.p28;
  o.rsc_floor = d.rsc_floor;
  o.pmc_num = d.pmc_num;
  o.pms_num = d.pms_num;
  o.inv_raa_rac = MathFast::rcp(d.raa_rac);
  o.inv_raa_ras = MathFast::rcp(d.raa_ras);
  o.inv_rac = MathFast::rcp(d.rac);
  o.inv_ras = MathFast::rcp(d.ras);
  o.Ra = d.Ra;
  o.dg_ras = d.dg_ras;
  o.dg_rac = d.dg_rac;
  o.rhocp = d.rhocp;
  o.raa_inv_rhocp = d.raa * MathFast::rcp(d.rhocp);
  o.lec_a = d.lec_a;
  o.les_a = d.les_a;
  o.k_lamb = 1.0E3f * MathFast::rcp(d.rhow_lamb);
  o.litter10 = d.litter10;
  o.litter1000 = d.litter1000;
  o.canopy_on = d.canopy_on;
}

__device__ __forceinline__ float fast_pow(float a, float b) { return MathFast::ex2(b * MathFast::lg2(a)); }

/* specific yield of layer I (0-based i) at water-table depth zwtmm, :963-965 */
template <class C>
__device__ __forceinline__ float fast_specific_yield(const C& c, int i, float zwtmm) {
  /* -1/bsw == (1 - 1/bsw) - 1, the table's e1 */
  const float s_y = c.ths(i) * (1.0f - fast_pow(fmaf(zwtmm, c.inv_npsi(i), 1.0f), c.e1(i) - 1.0f));
  return fmaxf(s_y, 0.02f);
}

/* Control flow of the sub-step.  One warp per scheduler (small shards) cannot hide a stall
 * behind another warp, and ptxas schedules inside basic blocks only, so the step is laid out as
 * TWO large basic blocks (measured budget: profiles/r02/cycle_budget_*.txt):
 *   block A  everything up to the end of the forward sweep, branch-free: the serial chains of
 *            the energy balance and of the aquifer layer run behind the 80 MUFU operations of
 *            the eight layers (8 cycles each for a lone warp) instead of in front of them;
 *   block B  the tail, ONE branch on `deep`: a straight-line tail for a water table below the
 *            column (the common case), the data-dependent Drainage code otherwise; each holds its
 *            own copy of the back substitution so that it overlaps the water-table chain;
 *   then     the clamp, the repair trigger as two max/min trees, and the water balance computed
 *            as if no repair were needed (the rarely taken repair branch redoes it).
 * Same operations on the same operands as before the re-ordering. */
/* sub-step variants: kStepThroughput is hydrology_step_fast_tp (h9_physics_fast_tp.cuh) */
enum : int { kStepThroughput = 0, kStepAllDeep = 1, kStepGeneral = 2 };

/* zi, zi/1000 and zc for layer indices that are only known at run time (the Drainage code).
 * One copy per block in shared memory: 30 floats in 30 different banks, so lanes that ask for
 * different layers are served in one pass.  The same lookups through the constant bank replay
 * once per distinct index -- measured on the general straight-line tail, where all 32 lanes
 * ask: 2,900 cycles per sub-step against 1,640 for the all-deep step. */
constexpr int kGeoDynFloats = 32;
struct GeoDyn {
  const float* p;
  __device__ __forceinline__ float zi(int i) const { H9_ASSERT(i >= 0 && i <= 9); return p[i]; }
  __device__ __forceinline__ float zim(int i) const { H9_ASSERT(i >= 0 && i <= 9); return p[10 + i]; }
  __device__ __forceinline__ float zc(int i) const { H9_ASSERT(i >= 0 && i <= 9); return p[20 + i]; }
  /* every thread of the block calls it before the first use (it ends with a barrier) */
  static __device__ __forceinline__ void fill(float* sh, const Geo& g) {
    if (threadIdx.x < 10) {
      sh[threadIdx.x] = g.zi[threadIdx.x];
      sh[10 + threadIdx.x] = g.zim[threadIdx.x];
      sh[20 + threadIdx.x] = g.zc[threadIdx.x];
    }
    __syncthreads();
  }
};

template <int MODE, class C>
__device__ __forceinline__ uint32_t hydrology_step_fast(const Geo& g, const GeoDyn& gd, const C& c,
                                                        const DayFast& d, FastState& s,
                                                        StepOut& o H9_TICKS_PARAM) {
  constexpr float kLog2e = 1.4426950408889634f;
  uint32_t fault = 0;
  const float dt = g.dt, rdt = g.rdt;
  float theta[NL];
  H9_TICK_START();

  /* ------------------------------ block A ------------------------------ */
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

  /* rsc :283-295, rss :325-331 (both branches of rss evaluated, then selected: no jump) */
  float rsc = (d.canopy_on && beta > 0.0f) ? d.rsc_num * MathFast::rcp(d.rsc_den0 * beta) : 1.0E6f;
  rsc = fmaxf(rsc, d.rsc_floor);
  const float rss_dry = d.litter10 * MathFast::ex2((35.63f * kLog2e) * (0.15f - theta[0]));
  const float rss_wet = fmaf(d.litter1000, 1.0f - theta[0] * c.inv_ths(0), 10.0f);
  const float rss = (theta[0] <= 0.15f) ? rss_dry : rss_wet;

  /* two-source Penman-Monteith :344-389 */
  const float dPMc = fmaf(d.gamma, fmaf(rsc, d.inv_raa_rac, 1.0f), d.desatdT);
  const float dPMs = fmaf(d.gamma, fmaf(rss, d.inv_raa_ras, 1.0f), d.desatdT);
  const float Rs = fmaf(d.gamma, rss, d.dg_ras);
  const float Rc = fmaf(d.gamma, rsc, d.dg_rac);
  /* Cc = 1/(1+Rc*Ra/(Rs*(Rc+Ra))) and Cs = 1/(1+Rs*Ra/(Rc*(Rs+Ra))) over the common
   * denominator Rs*Rc + Rs*Ra + Rc*Ra; three reciprocals from one MUFU */
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
  const float evap_max1 =
      fmaxf(0.0f, fmaf(-qflx_tran_veg_col, rr[0], g.dz[1] * (theta[0] - kWatmin) * rdt));
  qflx_evap_grnd = fminf(evap_max1, qflx_evap_grnd);

  /* Infiltration :426-478 */
  const float qflx_in_soil = (d.forc_rain - qflx_surf) - qflx_evap_grnd;
  const float qinmax = (1.0f - fsat) * c.hksmin3();
  const float qflx_infl_excess = fmaxf(0.0f, qflx_in_soil - qinmax);
  const float qflx_infl = qflx_in_soil - qflx_infl_excess;
  qflx_surf += qflx_infl_excess;

  /* SoilWater :492-508.  jwt == NL  <=>  zwt > zi(8)/1000 (the interfaces increase); the index
   * itself is only needed by the tail of a cell whose water table is inside the column */
  float zwtmm = 1000.0f * s.zwt;
  const bool deep = !(s.zwt <= g.zim[NL]);

  /* equilibrium profile :517-590, branch-free.  With B0 = (-psi+zwtmm-zi(I-1))/(-psi) and
   * BI = (-psi+zwtmm-zi(I))/(-psi), the three cases of the reference are one formula:
   *   vol = [ A*(max(BI,1)**e1 - max(B0,1)**e1) + ths*max(zi(I)-zwtmm,0) ] / dz,  A = psi*ths/e1
   * below the layer both bases are >= 1 and the last term vanishes (:548-558); inside it
   * BI < 1 so the first power is tempi = 1 (:532-540); above it both powers are 1 and
   * (zi(I)-zwtmm)/dz >= 1 so the clamp to theta_s returns theta_s (:523). */
  float zq[NL + 1];
  float hk[NL], dhkdw[NL], dsmpdw[NL], snode[NL];
  float q[NL], qa[NL], qb[NL]; /* per interface i (below layer i): flux and its two derivatives */
  float dwat2[NL + 1], gam[NL + 1];
  float bet = 1.0f, rbet = 1.0f, minpiv = 3.0e38f;
  const float dz9 = deep ? (zwtmm - g.zc[NL]) : g.dz[NL];
  const float4 rc = c.rootr4(0), rd = c.rootr4(1); /* re-read: cheaper than eight live registers */
  const float rs[NL] = {rc.x, rc.y, rc.z, rc.w, rd.x, rd.y, rd.z, rd.w};
  /* the aquifer's specific yield and its reciprocal (:925-940): known from the old zwt */
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
    fault |= (bet == 0.0f) ? FAULT_PIVOT1 : 0u;
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
  H9_TICK_FINE(0); /* energy balance and the aquifer layer, issued in front of the layers */
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
  H9_TICK_FINE(1); /* eight layers and rows 1-7 */
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
  fault |= (minpiv == 0.0f) ? FAULT_PIVOT2 : 0u; /* BET == 0 in some row :818 */

  /* for a step that may hold cells with the water table inside the column: the inputs of
   * their tail that do not depend on the solve -- the index jwt :499-508 (as a count: the
   * interfaces increase), the Darcy recharge :866-895 (reads the NEW smp, :880) and the specific
   * yield of the layer holding the table -- still inside block A, behind the sweep.  Cells
   * with a deep table compute them on layer 8's constants and drop them. */
  int jc = NL - 1;
  float qcharge_early = 0.0f, sy_first = 0.02f, rsy_first = 50.0f;
  if (MODE == kStepGeneral) {
    /* (as an add tree: the count sits on the dependent chain of the tail) */
    const int cnt = (((s.zwt > g.zim[1]) + (s.zwt > g.zim[2])) + ((s.zwt > g.zim[3]) + (s.zwt > g.zim[4]))) +
                    (((s.zwt > g.zim[5]) + (s.zwt > g.zim[6])) + ((s.zwt > g.zim[7]) + (s.zwt > g.zim[8])));
    jc = (cnt < NL) ? cnt : NL - 1;
    const int jm = (jc > 1 ? jc : 1) - 1;
    const float th_j = pick<NL>(theta, jc);
    const float s1 = fminf(1.0f, fmaxf(th_j * c.inv_ths(jc), 0.01f));
    const float ka = c.hks(jc) * fast_pow(s1, fmaf(2.0f, c.bsw(jc), 3.0f));
    const float smp1 = fmaxf(kSmpmin, pick<NL>(s.smp, jm));
    float zq_j = zq[0];
#pragma unroll
    for (int k = 1; k < NL; ++k)
      if (jm == k) zq_j = zq[k];
    const float wh = smp1 - zq_j;
    const float denom = (jc == 0) ? (zwtmm + 1.0f) : (zwtmm - gd.zc(jc)) * 2.0f;
    qcharge_early = ka * wh * MathFast::rcp(denom); /* -ka*(0 - wh)/denom */
    qcharge_early = fminf(g.q10_hi, fmaxf(g.q10_lo, qcharge_early));
    sy_first = fast_specific_yield(c, jc, zwtmm); /* layer jwt+1 */
    rsy_first = MathFast::rcp(sy_first);
  }
  H9_TICK(2); /* aquifer node, rows 8-9 */

  /* ------------------------------ block B ------------------------------ */
  /* back substitution :828-830 and state update :845-850 (only dwat2(9) feeds Recharge and
   * Drainage, so it is scheduled beside the water-table chain of whichever tail follows) */
  auto back_substitute = [&]() {
#pragma unroll
    for (int i = NL - 1; i >= 0; --i) dwat2[i] = fmaf(-gam[i + 1], dwat2[i + 1], dwat2[i]);
#pragma unroll
    for (int i = 0; i < NL; ++i) s.h2o[i] = fmaf(dwat2[i], g.dz[i + 1], s.h2o[i]);
  };
  float rsub_top;
  int jwt = NL;
  float zwt_j;
  float imb, rnf_inc;
  bool needfix;
  const float rnf_sum_in = s.rnf_sum;
  uint32_t fault_bal = 0, fault_tail = 0;
  /* balance :1221-1244 and the runoff sum :1282-1283 */
  auto balance = [&](float qflx_rsub_sat) {
    const float w1 =
        fmaf((qflx_surf + qflx_evap_grnd + qflx_tran_veg_col) + rsub_top + qflx_rsub_sat, dt, s.wa) +
        (((s.h2o[0] + s.h2o[1]) + (s.h2o[2] + s.h2o[3])) + ((s.h2o[4] + s.h2o[5]) + (s.h2o[6] + s.h2o[7])));
    imb = w1 - w0;
    fault_bal = (!(fabsf(imb) <= 0.1f)) ? FAULT_IMBAL : 0u;
    const float r1 = qflx_surf * dt, r2 = rsub_top * dt;
    rnf_inc = r1 + r2;
    s.rnf_sum = fmaf(rsub_top, dt, fmaf(qflx_surf, dt, rnf_sum_in));
  };
  auto excess_tree = [&]() { /* max over the layers of (water - capacity), :1131-1148 */
    float ov[NL];
    ov[0] = s.h2o[0] - fmaxf(0.0f, c.ths(0) * g.dz[1]);
#pragma unroll
    for (int i = 1; i < NL; ++i) ov[i] = fmaf(-fmaxf(0.01f, c.ths(i)), g.dz[i + 1], s.h2o[i]);
    return fmaxf(fmaxf(fmaxf(ov[0], ov[1]), fmaxf(ov[2], ov[3])),
                 fmaxf(fmaxf(ov[4], ov[5]), fmaxf(ov[6], ov[7])));
  };
  auto lowest_tree = [&]() {
    return fminf(fminf(fminf(s.h2o[0], s.h2o[1]), fminf(s.h2o[2], s.h2o[3])),
                 fminf(fminf(s.h2o[4], s.h2o[5]), fminf(s.h2o[6], s.h2o[7])));
  };
  /* End of a tail whose water table is below the column (jwt == NL): clamp :1122-1123, then the
   * dryness repair :1161-1174 itself, straight line.  Cells of dry regions sit at watmin and
   * need it EVERY sub-step (transpiration is not limited by the water of the layer, G18), and one
   * such cell makes its warp pay for a branch: measured 1,970 against 1,530 cycles per sub-step
   * (profiles/r02/cycle_budget_*.txt).  With jwt == NL the pass never touches zwt (:1166), so it
   * is seven dependent max/add/sub triples behind the back substitution, in the shadow of the
   * baseflow chain.  What stays behind the rare branch: the excess cascade :1131-1152 (it comes
   * first in the reference, so the branch restarts from the saved column) and the upward search
   * of the bottom layer :1181-1198. */
  float h_pre[NL];
  auto finish_deep = [&]() {
    s.zwt = fminf(80.0f, fmaxf(0.0f, zwt_j));
    const float over = excess_tree();
#pragma unroll
    for (int i = 0; i < NL; ++i) h_pre[i] = s.h2o[i];
#pragma unroll
    for (int i = 0; i < NL - 1; ++i) {
      const float xs = fmaxf(kWatmin - s.h2o[i], 0.0f); /* > 0 exactly when h2o < watmin */
      s.h2o[i] += xs;
      s.h2o[i + 1] -= xs;
    }
    needfix = (over > 0.0f || s.h2o[NL - 1] < kWatmin);
    balance(0.0f);
  };
  /* end of a tail in general: clamp, the repair trigger as two max / min trees, and the balance
   * as if no repair were needed */
  auto finish = [&]() {
    s.zwt = fminf(80.0f, fmaxf(0.0f, zwt_j));
    needfix = (excess_tree() > 0.0f || lowest_tree() < kWatmin);
    balance(0.0f);
  };
  /* excess cascade :1131-1152 and dryness repair :1161-1205: both are no-ops unless some layer
   * is above its capacity / below watmin, so ONE rarely taken branch guards them (trigger and
   * water balance were computed in the tail as if no repair were needed; the branch redoes the
   * balance).  The repair's trigger is taken before the cascade: the cascade only lowers layers
   * to their capacity (>= 0.01*dz > watmin) and raises others, so it cannot create a layer below
   * watmin, and inside the branch the repair works from the current values anyway. */
  auto repair = [&]() {
#ifdef H9_CYCLE_BUDGET
    tk.acc[kTickSegs - 1] += 1u; /* how often this thread takes the branch */
#endif
    float qflx_rsub_sat = 0.0f;
    float xs = 0.0f;
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
    s.h2o[NL - 1] += xs;  /* :1205 */
    rsub_top -= xs * rdt; /* :1211 */
    return qflx_rsub_sat;
  };
  /* Recharge :896-904, Drainage :946-951 and baseflow :1048-1058 of a water table below the
   * column, straight-line.  jwt stays NL: it is not recomputed on this path (G10) */
  auto deep_tail = [&]() {
    const float qcharge = dwat2[NL] * dz9 * rdt;
    s.wa = fmaf(qcharge, dt, s.wa);
    s.zwt = fmaf(-(qcharge * dt * 0.001f), rrous_early, s.zwt);
    zwtmm = 1000.0f * s.zwt; /* :1015 */
    rsub_top = 5.5E-3f * MathFast::ex2((-kFff * kLog2e) * s.zwt);
    const float rous = fast_specific_yield(c, NL - 1, zwtmm);
    s.wa = fmaf(-rsub_top, dt, s.wa);
    s.zwt = fmaf(rsub_top * dt * 0.001f, MathFast::rcp(rous), s.zwt);
    s.h2o[NL - 1] += fmaxf(0.0f, s.wa - 5000.0f);
    s.wa = fminf(s.wa, 5000.0f);
    zwt_j = s.zwt;
  };
  /* the data-dependent tail of a water table inside the column, as the reference loops it */
  auto shallow_tail = [&](float qcharge, float sy1) {
    /* Drainage :953-1009; zwtmm stays the stale value of :492 inside the loops (G9) */
    const int jfirst = jwt + 1;
    float qcharge_tot = qcharge * dt;
    if (qcharge_tot > 0.0f) { /* rising, layers jwt+1 .. 1 */
      for (int I = jwt + 1; I >= 1; --I) {
        H9_ASSERT(I >= 1 && I <= NL);
        const float s_y = (I == jfirst) ? sy1 : fast_specific_yield(c, I - 1, zwtmm);
        const float ql = fmaxf(fminf(qcharge_tot, s_y * (zwtmm - gd.zi(I - 1))), 0.0f);
        s.zwt -= ql * MathFast::rcp(s_y) * 0.001f; /* s_y >= 0.02 > 0 */
        qcharge_tot -= ql;
        if (qcharge_tot <= 0.0f) break;
      }
    } else { /* deepening, layers jwt+1 .. 8 */
      for (int I = jwt + 1; I <= NL; ++I) {
        H9_ASSERT(I >= 1 && I <= NL);
        const float s_y = (I == jfirst) ? sy1 : fast_specific_yield(c, I - 1, zwtmm);
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
    /* baseflow :1024-1118 */
    rsub_top = 5.5E-3f * MathFast::ex2((-kFff * kLog2e) * s.zwt);
    const float rous = fast_specific_yield(c, NL - 1, zwtmm);
    if (jwt == NL) { /* the recharge pushed the table below the column */
      s.wa = fmaf(-rsub_top, dt, s.wa);
      s.zwt = fmaf(rsub_top * dt * 0.001f, MathFast::rcp(rous), s.zwt);
      s.h2o[NL - 1] += fmaxf(0.0f, s.wa - 5000.0f);
      s.wa = fminf(s.wa, 5000.0f);
    } else {
      float rsub_top_tot = -rsub_top * dt;
      if (rsub_top_tot > 0.0f) {
        fault_tail |= FAULT_RSUB;
      } else {
        for (int I = jwt + 1; I <= NL; ++I) {
          H9_ASSERT(I >= 1 && I <= NL);
          const float s_y = fast_specific_yield(c, I - 1, zwtmm);
          const float rl = fminf(fmaxf(rsub_top_tot, -(s_y * (gd.zi(I) - zwtmm))), 0.0f);
#pragma unroll
          for (int k = 0; k < NL; ++k)
            s.h2o[k] = (k == I - 1) ? s.h2o[k] + rl : s.h2o[k]; /* a select, not a jump table */
          rsub_top_tot -= rl;
          if (rsub_top_tot >= 0.0f) {
            s.zwt -= rl * MathFast::rcp(s_y) * 0.001f;
            break;
          }
          s.zwt = gd.zim(I);
        }
        s.zwt -= rsub_top_tot * 0.001f * MathFast::rcp(rous); /* residual, unconditional :1100-1101 */
        s.wa += rsub_top_tot;
      }
      jwt = -1; /* :1110-1116, evaluated from zwt_j only where it is read (dryness repair, output) */
    }
    zwt_j = s.zwt;
  };
  if (MODE == kStepAllDeep) {
    /* every cell of the warp has its table below the column: no branch but the rare repair */
    back_substitute();
    deep_tail();
    finish_deep();
    if (needfix) {
#pragma unroll
      for (int i = 0; i < NL; ++i) s.h2o[i] = h_pre[i]; /* the cascade comes first */
      balance(repair());
    }
  } else {
    /* General straight-line tail: the FIRST iteration of each Drainage loop (:961-1009,
     * :1075-1098) with selects -- almost always the only one -- for cells with the table inside
     * the column, the deep formulas for the others, one instruction stream for the warp.  A cell
     * whose loop would go on, or that needs the cascade, the repair with its water-table
     * move, or the bottom layer's search, restarts the tail from the saved state on the exact
     * looping code (rare). */
    back_substitute();
    const float zwt0 = s.zwt, wa0 = s.wa;
    float h_bs[NL];
#pragma unroll
    for (int i = 0; i < NL; ++i) h_bs[i] = s.h2o[i];
    const float qcharge = deep ? dwat2[NL] * dz9 * rdt : qcharge_early;
    const float qtot = qcharge * dt;
    H9_ASSERT(jc >= 0 && jc + 1 <= NL);
    /* Drainage, first layer (jc = jwt, the layer holding the table) */
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
    /* baseflow, first layer below the new table */
    const float zwtmm1 = 1000.0f * zwt1; /* :1015 */
    rsub_top = 5.5E-3f * MathFast::ex2((-kFff * kLog2e) * zwt1);
    const float rous = fast_specific_yield(c, NL - 1, zwtmm1);
    const float rrous = MathFast::rcp(rous);
    const int cnt2 = (((zwt1 > g.zim[1]) + (zwt1 > g.zim[2])) + ((zwt1 > g.zim[3]) + (zwt1 > g.zim[4]))) +
                     (((zwt1 > g.zim[5]) + (zwt1 > g.zim[6])) + ((zwt1 > g.zim[7]) + (zwt1 > g.zim[8])));
    const bool isA = deep || cnt2 == NL; /* table below the column: :1048-1058 */
    const int j2c = (cnt2 < NL) ? cnt2 : NL - 1;
    H9_ASSERT(j2c >= 0 && j2c + 1 <= NL);
    const float sy2 = fast_specific_yield(c, j2c, zwtmm1);
    const float rtot = -rsub_top * dt;
    const float rl = fminf(fmaxf(rtot, -(sy2 * (gd.zi(j2c + 1) - zwtmm1))), 0.0f);
    const float rrem = rtot - rl;
    const bool done2 = rrem >= 0.0f;
    const float zwt_b = done2 ? zwt1 - rl * MathFast::rcp(sy2) * 0.001f : gd.zim(j2c + 1);
    const bool more2 = !isA && ((!done2 && j2c + 2 <= NL) || rtot > 0.0f);
    const float zwt_c = zwt_b - rrem * 0.001f * rrous; /* residual, unconditional :1100-1101 */
    const float wa_c = wa1 + rrem;
    const float wa_a = fmaf(-rsub_top, dt, wa1);
    const float zwt_a = fmaf(rsub_top * dt * 0.001f, rrous, zwt1);
    zwt_j = isA ? zwt_a : zwt_c;
    jwt = isA ? NL : -1; /* :1110-1116 only where it is read (the output of the one-step entry) */
    const float wa2 = isA ? wa_a : wa_c;
#pragma unroll
    for (int k = 0; k < NL; ++k) s.h2o[k] = (!isA && k == j2c) ? s.h2o[k] + rl : s.h2o[k];
    s.h2o[NL - 1] += isA ? fmaxf(0.0f, wa2 - 5000.0f) : 0.0f;
    s.wa = isA ? fminf(wa2, 5000.0f) : wa2;
    /* clamp :1122-1123, triggers, the dryness pass where it cannot move the table (jwt == NL,
     * :1166), balance.  A straight-line cascade :1131-1152 was measured too: it lengthens the
     * dependent chain of every warp on this tail by ~100 cycles and the cells that need it are
     * too few to pay for that (profiles/r02/README.md), so it stays behind the rare restart. */
    s.zwt = fminf(80.0f, fmaxf(0.0f, zwt_j));
    const float over = excess_tree();
    const float lowest = lowest_tree();
#pragma unroll
    for (int i = 0; i < NL - 1; ++i) {
      const float xs = isA ? fmaxf(kWatmin - s.h2o[i], 0.0f) : 0.0f;
      s.h2o[i] += xs;
      s.h2o[i + 1] -= xs;
    }
    const bool fix = over > 0.0f || (isA ? (s.h2o[NL - 1] < kWatmin) : (lowest < kWatmin));
    balance(0.0f);
#ifdef H9_CYCLE_BUDGET
    tk.acc[kTickSegs - 2] += 1u;                                  /* general-mode sub-steps */
    tk.acc[kTickSegs - 3] += (more1 && !deep) ? 1u : 0u;
    tk.acc[kTickSegs - 4] += more2 ? 1u : 0u;
    tk.acc[kTickSegs - 5] += fix ? 1u : 0u;
#endif
    if ((more1 && !deep) || more2 || fix) { /* rare: the exact looping code from the saved state */
      s.zwt = zwt0;
      s.wa = wa0;
#pragma unroll
      for (int i = 0; i < NL; ++i) s.h2o[i] = h_bs[i];
      zwtmm = 1000.0f * zwt0;
      if (deep) {
        deep_tail();
      } else {
        jwt = jc;
        shallow_tail(qcharge_early, sy_first);
      }
      finish();
      if (needfix) balance(repair());
    }
  }
  fault |= fault_tail;
  H9_TICK(3); /* tail: recharge, drainage, baseflow, back substitution, trigger, balance, rare repair */

  o.qflx_tran_veg_col = qflx_tran_veg_col;
  o.qflx_evap_grnd = qflx_evap_grnd;
  o.rnf_inc = rnf_inc;
  o.imbalance = imb;
  o.jwt = (jwt < 0) ? find_jwt(g, zwt_j) : jwt; /* dead code where the caller ignores it (K3) */
  return fault | fault_bal;
}

} /* namespace h9 */
#endif
