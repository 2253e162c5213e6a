/*
 * h9_physics.h -- per-cell physics of the HYBRID9 hot path, written for one
 * GPU thread per land cell with every per-layer array held in registers.
 *
 * Restates (citations relative to /root/reference/SOURCE):
 *   HYBRID9.f90:168-184     forcing derivation          -> day_setup()
 *   HYDROLOGY.f90:141-1283  one sub-step of one cell    -> hydrology_step()
 *   GROW.f90:55-201         one day of one cell         -> grow_day()
 *   INIT.f90:791-797        root profile                -> root_profile()
 *
 * Structure differs from the Fortran on purpose:
 *   - everything that is constant over the NISURF sub-steps of a day (all the
 *     meteorology, the LAI-only resistances, the numerators of the two
 *     Penman-Monteith terms) is evaluated once per day in day_setup(); the
 *     sub-expressions are the reference's own, so the values are bit-identical
 *     to evaluating them every sub-step;
 *   - data-dependent layer loops (water-table search, Drainage) are unrolled
 *     and predicated so that no array is indexed dynamically (no local memory);
 *   - module scratch (zc(9), dz(9), smp, theta ...) is thread-local / per cell.
 *
 * The arithmetic policy M supplies pow/exp/log/div: MathExact keeps the
 * reference's operation order with IEEE division (its translation unit is
 * built with -fmad=false) and computes pow/exp/log with portable double-precision
 * kernels that give the same bits on the GPU and on a host build; MathFast uses
 * MUFU ex2/lg2/rcp (its sub-step lives in h9_physics_fast.cuh with explicit FMAs).
 *
 * The header is plain C++ so that tests/twin can compile the very same source
 * for the host and diff it against the oracle without a GPU.
 */
#ifndef H9_PHYSICS_H
#define H9_PHYSICS_H

#include <float.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "h9_exact_tables.h"

#if defined(__CUDACC__)
#define H9_HD __host__ __device__ __forceinline__
/* the exact mode's pow / exp / log are CALLED, not inlined, on the device: inlined into the ~50
 * sites of a sub-step they made a 20,000-instruction kernel that spent its time missing the
 * instruction cache (profiles/r02/README.md section 11) */
#define H9_HD_CALL static __host__ __device__ __noinline__
#define H9_UNROLL _Pragma("unroll")
#else
#define H9_HD inline
#define H9_HD_CALL inline
#define H9_UNROLL
#endif

/* Device-side index assertions: compiled in only by -DH9_BOUNDS_CHECK (tools/bounds_check.sh,
 * the stand-in for compute-sanitizer's memcheck on this GPU pool); a violation traps the kernel
 * and surfaces as cudaErrorAssert through the C ABI. */
#if defined(H9_BOUNDS_CHECK) && defined(__CUDACC__)
#include <assert.h>
#define H9_ASSERT(x) assert(x)
#else
#define H9_ASSERT(x) ((void)0)
#endif

namespace h9 {

constexpr int NL = 8; /* nsoil_layers_max, SHARED.f90:294 */

/* fault bits, same values as include/h9gpu.h */
constexpr uint32_t FAULT_PIVOT1 = 1u, FAULT_PIVOT2 = 2u, FAULT_RSUB = 4u, FAULT_IMBAL = 8u;

/* constants of SHARED.f90:308-367 and HYDROLOGY.f90:35,135 (float, as default REAL) */
constexpr float kRhow = 1000.0f;
constexpr float kGasc = 8.314510f;
constexpr float kMair = 28.9655f;
constexpr float kMwat = 18.015f;
constexpr float kRgas = 1000.0f * kGasc / kMair;
constexpr float kDeltx = 1.0f / (kMwat / kMair) - 1.0f;
constexpr float kStbo = 5.67E-8f;
constexpr float kTf = 273.16f;
constexpr float kSmpmin = -1.0E8f;
constexpr float kCp = 1010.0f;
constexpr float kWatmin = 0.01f;
constexpr float kSla = 23.0E-3f; /* INIT.f90:154 */
constexpr float kFff = 2.5f;     /* 1/(1/2.5) in float is exactly 2.5, HYDROLOGY.f90:182-188 */

/* geometry, INIT.f90:214,252-257; index 0 of dz/zc unused, zi is zi(0:9) */
struct Geo {
  float zi[10];
  float dz[10];
  float zc[10];
  float zim[10];  /* zi(I)/1000, the right-hand side of the jwt tests, HYDROLOGY.f90:504 */
  float dzw[10];  /* dz(I)*rhow/1.0E3, HYDROLOGY.f90:148,1234 */
  float dzdt[10]; /* dz(I)/dt, HYDROLOGY.f90:674,700,727 */
  float zi10[10]; /* zi(I)/10, GROW.f90:180-181 */
  /* reciprocals used by the fast mode only */
  float rdzw[10]; /* 1/dzw(I) */
  float rdzl[10]; /* 1/(zi(I)-zi(I-1)) */
  float rden[10]; /* 1/(zc(I+1)-zc(I)), I = 1..7 */
  float zhr[10];  /* zi(I)/(zi(I)-zi(I-1)): (zi(I)-zwtmm)/dz(I) == zhr(I) - zwtmm*rdzl(I) */
  float kbeta[10]; /* 1 - zc(I)/150000: the beta term 1-(smp-zc)/(-150000) == smp/150000 + kbeta */
  float rdt;      /* 1/dt */
  float dt;
  float q10_lo, q10_hi; /* -10/dt, 10/dt, HYDROLOGY.f90:894-895 */
  int nisurf;
};

/* host-side fill of Geo from zi(0:9) and NISURF; float IEEE ops, as INIT does */
inline void geo_init(Geo& g, const float zi[10], int nisurf) {
  g.nisurf = nisurf;
  g.dt = 86400.0f / (float)nisurf; /* INIT.f90:214 */
  for (int I = 0; I < 10; ++I) g.zi[I] = zi[I];
  g.dz[0] = g.zc[0] = 0.0f;
  for (int I = 1; I <= 9; ++I) g.dz[I] = g.zi[I] - g.zi[I - 1];        /* INIT.f90:253 */
  for (int I = 1; I <= 9; ++I) g.zc[I] = g.zi[I] - g.dz[I] / 2.0f;     /* INIT.f90:256 */
  for (int I = 0; I < 10; ++I) {
    g.zim[I] = g.zi[I] / 1000.0f;
    g.dzw[I] = g.dz[I] * kRhow / 1.0E3f;
    g.dzdt[I] = g.dz[I] / g.dt;
    g.zi10[I] = g.zi[I] / 10.0f;
    g.rdzw[I] = I >= 1 ? 1.0f / g.dzw[I] : 0.0f;
    g.rdzl[I] = I >= 1 ? 1.0f / g.dz[I] : 0.0f;
    g.rden[I] = (I >= 1 && I <= 8) ? 1.0f / (g.zc[I + 1] - g.zc[I]) : 0.0f;
    g.zhr[I] = I >= 1 ? g.zi[I] / g.dz[I] : 0.0f;
    g.kbeta[I] = 1.0f - g.zc[I] / 150000.0f; /* 1 + (smp - zc)/150000 == smp/150000 + kbeta */
  }
  g.rdt = 1.0f / g.dt;
  g.q10_lo = -10.0f / g.dt;
  g.q10_hi = 10.0f / g.dt;
}

struct Params { /* SHARED.f90:398-446 */
  float theta_s[NL], hksat[NL], bsw[NL], psi_s[NL];
  float fmax;
};

struct State { /* SHARED.f90:30-101,198,459-472 */
  float h2o[NL];   /* h2osoi_liq */
  float smp[NL];   /* promoted from module scratch to per-cell state */
  float rootr[NL]; /* rootr_col(1:8); element 9 is always 0 */
  float zwt, wa, lai, lai_litter;
  float plant_mass, plant_foliage_mass, plant_length, rdepth;
  float rnf_sum; /* SHARED.f90:134, reset by the driver each year */
  int nplants;
};

struct Forcing { /* one day of PGF forcing for the cell, READ_PGF.f90:24-109 */
  float tas, rlds, rsds, huss, ps, pr, rhs;
};

struct StepOut {
  float qflx_tran_veg_col, qflx_evap_grnd, rnf_inc, imbalance;
  int jwt;
};

/* ---- arithmetic policies ------------------------------------------------ */

/* Portable transcendental kernels for the exact mode: IEEE double +,*,/ and fma
 * only (no libm, no vendor intrinsics), so that the CUDA kernel and a host
 * build of this header produce the SAME BITS.  Accuracy: relative error below
 * 1e-14 before the final rounding to float, i.e. the float result is correctly
 * rounded except within ~1e-7 ulp of a rounding boundary. */
H9_HD double h9_bits_to_double(uint64_t u) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)u);
#else
  double d;
  memcpy(&d, &u, sizeof(d));
  return d;
#endif
}
H9_HD uint64_t h9_double_to_bits(double d) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(d);
#else
  uint64_t u;
  memcpy(&u, &d, sizeof(u));
  return u;
#endif
}

H9_HD int h9_hi_word(double d) {
#if defined(__CUDA_ARCH__)
  return __double2hiint(d);
#else
  return (int)(uint32_t)(h9_double_to_bits(d) >> 32);
#endif
}
H9_HD int h9_lo_word(double d) {
#if defined(__CUDA_ARCH__)
  return __double2loint(d);
#else
  return (int)(uint32_t)(h9_double_to_bits(d) & 0xFFFFFFFFull);
#endif
}
H9_HD double h9_from_words(int hi, int lo) {
#if defined(__CUDA_ARCH__)
  return __hiloint2double(hi, lo);
#else
  return h9_bits_to_double(((uint64_t)(uint32_t)hi << 32) | (uint64_t)(uint32_t)lo);
#endif
}

/* Tables of the two kernels (tools/gen_exact_tables.py; hexadecimal literals: the same bits in
 * every build).  On the device the per-lane lookups read global memory (L1-resident; a constant
 * bank would serialise the 32 different indices of a warp) and the polynomial coefficients sit
 * in a constant bank, where an FP64 instruction takes them as an operand: as literals each
 * cost two uniform moves, a fifth of the exact kernel's executed instructions. */
struct H9Log2Row {
  double inv_c, log2_c;
};
#define H9_TAB_ROW(a, b) {a, b},
#define H9_TAB_ONE(a) a,
#define H9_COEF_LIST(X)                                                                          \
  X(-1.0 / 8.0) X(1.0 / 7.0) X(-1.0 / 6.0) X(1.0 / 5.0) X(-1.0 / 4.0) X(1.0 / 3.0) X(-1.0 / 2.0) \
  X(1.4426950408889634) X(6755399441055744.0) X(-1.0 / 32.0) X(0.6931471805599453)               \
  X(1.0 / 720.0) X(1.0 / 120.0) X(1.0 / 24.0) X(1.0 / 6.0)
enum {
  kH9cL8, kH9cL7, kH9cL6, kH9cL5, kH9cL4, kH9cL3, kH9cL2, kH9cLog2e, kH9cBig, kH9cM32, kH9cLn2,
  kH9cE6, kH9cE5, kH9cE4, kH9cE3
};
static const H9Log2Row kH9Log2_h[47] = {H9_LOG2_TABLE(H9_TAB_ROW)};
static const double kH9Exp2_h[32] = {H9_EXP2_TABLE(H9_TAB_ONE)};
static const double kH9Coef_h[15] = {H9_COEF_LIST(H9_TAB_ONE)};
#if defined(__CUDACC__)
static __device__ const H9Log2Row kH9Log2_d[47] = {H9_LOG2_TABLE(H9_TAB_ROW)};
static __device__ const double kH9Exp2_d[32] = {H9_EXP2_TABLE(H9_TAB_ONE)};
static __constant__ double kH9Coef_d[15] = {H9_COEF_LIST(H9_TAB_ONE)};
#endif
#if defined(__CUDA_ARCH__)
#define H9_TAB(name) name##_d
#else
#define H9_TAB(name) name##_h
#endif
#define H9_K(i) H9_TAB(kH9Coef)[i]

/* log2(x) for finite x > 0 that is a normal double (every positive float is).
 * x = 2^e * m with m in [0.70710, 1.41421) (the split is made on the high word, at sqrt(2)
 * truncated to 20 bits); m = c (1 + r) with c = (45+i)/64 the nearest centre (c = 1 among them,
 * so that x near 1 keeps its relative accuracy), |r| <= 0.0112;
 * log2 x = e + log2 c + log2(e) * (r - r^2/2 + ... - r^8/8); the dropped term is below 3e-19.
 * No division, eight dependent FMAs. */
H9_HD double h9_log2_pos_inl(double x) {
  int hx = h9_hi_word(x) + (0x3FF00000 - 0x3FE6A09E);
  const int e = (hx >> 20) - 0x3FF;
  hx = (hx & 0x000FFFFF) + 0x3FE6A09E;
  const double m = h9_from_words(hx, h9_lo_word(x));
  /* round(64 m) in the low word of 64 m + 1.5 * 2^52 */
  const int i = h9_lo_word(fma(m, 64.0, H9_K(kH9cBig))) - 45; /* 0..46 */
  const H9Log2Row row = H9_TAB(kH9Log2)[i];
  const double r = fma(m, row.inv_c, -1.0);
  double p = H9_K(kH9cL8);
  p = fma(p, r, H9_K(kH9cL7));
  p = fma(p, r, H9_K(kH9cL6));
  p = fma(p, r, H9_K(kH9cL5));
  p = fma(p, r, H9_K(kH9cL4));
  p = fma(p, r, H9_K(kH9cL3));
  p = fma(p, r, H9_K(kH9cL2));
  p = fma(p, r, 1.0);
  const double lnm = p * r;
  return fma(lnm, H9_K(kH9cLog2e), row.log2_c + (double)e);
}

/* 2**y for any double y that is not a NaN (saturates far outside the float range).
 * y = k/32 + r with k = round(32 y), |r| <= 1/64; 2^y = 2^(k div 32) * 2^((k mod 32)/32) * e^(r ln 2),
 * the middle factor from the table with k div 32 added to its exponent field, the last one by
 * its Taylor series to degree 6 (the dropped term is below 4e-18). */
H9_HD double h9_exp2_inl(double y) {
  const int hy = h9_hi_word(y);
  if ((hy & 0x7FFFFFFF) >= 0x4072C000) y = hy < 0 ? -300.0 : 300.0; /* |y| >= 300, infinities */
  const double big = H9_K(kH9cBig); /* 1.5 * 2^52: round(32 y) lands in the low word */
  const double kb = fma(y, 32.0, big);
  const int ki = h9_lo_word(kb);
  const double k = kb - big;
  const double t = fma(k, H9_K(kH9cM32), y) * H9_K(kH9cLn2); /* |t| <= 0.0109 */
  double p = H9_K(kH9cE6);
  p = fma(p, t, H9_K(kH9cE5));
  p = fma(p, t, H9_K(kH9cE4));
  p = fma(p, t, H9_K(kH9cE3));
  p = fma(p, t, 0.5);
  p = fma(p, t, 1.0);
  p = fma(p, t, 1.0);
  const double tj = H9_TAB(kH9Exp2)[ki & 31];
  /* 2^((k mod 32)/32) * 2^(k div 32): exact, the exponent stays within [-301, 301] */
  return h9_from_words(h9_hi_word(tj) + ((ki >> 5) << 20), h9_lo_word(tj)) * p;
}

/* a**b with Fortran REAL semantics on the domain this path uses (a >= 0) */
H9_HD_CALL float h9_pow_f32(float a, float b) {
  if (!(a > 0.0f && a <= FLT_MAX && fabsf(b) <= FLT_MAX)) { /* one test on the common path */
    if (b == 0.0f || a == 1.0f) return 1.0f;
    if (!(a == a) || !(b == b)) return a + b; /* NaN */
    if (a < 0.0f) return NAN;                 /* negative base, real exponent */
    if (a == 0.0f) return b > 0.0f ? 0.0f : INFINITY;
    if (a > FLT_MAX) return b > 0.0f ? a : 0.0f; /* +inf */
    /* left: b infinite, 0 < a < inf, a != 1: the product below is the right infinity */
  }
  return (float)h9_exp2_inl((double)b * h9_log2_pos_inl((double)a));
}
H9_HD_CALL float h9_exp_f32(float a) {
  if (!(a == a)) return a;
  return (float)h9_exp2_inl((double)a * H9_K(kH9cLog2e));
}
H9_HD_CALL float h9_log_f32(float a) {
  if (!(a > 0.0f)) return a == 0.0f ? -INFINITY : NAN;
  if (a > FLT_MAX) return a;
  return (float)(h9_log2_pos_inl((double)a) * H9_K(kH9cLn2));
}

struct MathExact {
  static constexpr bool kFast = false;
  static H9_HD float pow(float a, float b) { return h9_pow_f32(a, b); }
  static H9_HD float exp(float a) { return h9_exp_f32(a); }
  static H9_HD float log(float a) { return h9_log_f32(a); }
  /* IEEE division stays inline: as a call it shrinks the kernel by a third and still runs
   * 11 % slower (profiles/r02/README.md section 11) */
  static H9_HD float div(float a, float b) { return a / b; }
};

#if !defined(__CUDACC__)
/* host-only policy with the C library's powf/expf/logf: what the oracle uses.
 * tests/twin builds the kernel source with it to diff the LOGIC against the
 * oracle bit for bit. */
struct MathLibm {
  static constexpr bool kFast = false;
  static inline float pow(float a, float b) { return powf(a, b); }
  static inline float exp(float a) { return expf(a); }
  static inline float log(float a) { return logf(a); }
  static inline float div(float a, float b) { return a / b; }
};
#endif

#if defined(__CUDACC__)
struct MathFast {
  static constexpr bool kFast = true;
  static __device__ __forceinline__ float lg2(float a) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
    return r;
  }
  static __device__ __forceinline__ float ex2(float a) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
    return r;
  }
  static __device__ __forceinline__ float rcp(float a) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
    return r;
  }
  /* every base on this path is positive and finite (DESIGN.md "pow") */
  static __device__ __forceinline__ float pow(float a, float b) { return ex2(b * lg2(a)); }
  static __device__ __forceinline__ float exp(float a) { return ex2(a * 1.4426950408889634f); }
  static __device__ __forceinline__ float log(float a) { return lg2(a) * 0.6931471805599453f; }
  static __device__ __forceinline__ float div(float a, float b) { return a * rcp(b); }
};
#endif

/* ---- per-day constants --------------------------------------------------- */

struct Day {
  float tas;       /* raw, GROW.f90:66 */
  float forc_rain; /* HYBRID9.f90:178 */
  float rain_dt;   /* forc_rain*dt, HYDROLOGY.f90:141 */
  float desatdT, gamma, dg, VDD; /* dg = desatdT+gamma */
  float rhocp;                   /* rho*cp */
  float rhow_lamb;               /* rhow*lamb, :388-389 */
  float rsc_num, lai2, p28, rsc_floor; /* :285-295 */
  float rac, raa, ras;
  float A;                    /* desatdT*(Rnet-G) */
  float pmc_num, pms_num;     /* numerators of :344-352 */
  float raa_rac, raa_ras;     /* raa+rac, raa+ras */
  float Ra;                   /* :356 */
  float dg_ras, dg_rac;       /* (desatdT+gamma)*ras, *rac :357-358 */
  float lec_a, les_a;         /* desatdT*(Rnet-Rnets), desatdT*(Rnets-G) :381-383 */
  float litter10, litter1000; /* 10+1000*LAI_litter ; 1000*LAI_litter :326-330 */
  bool canopy_on;             /* LAI>0 .AND. PAR>0, :283 */
};

template <class M>
H9_HD void day_setup(const Geo& g, const Forcing& f, float lai, float lai_litter, Day& d) {
  /* HYBRID9.f90:168-184 */
  const float tak = f.tas;
  const float t2 = f.tas * f.tas;
  const float Rnet = (0.92f * f.rsds + f.rlds) - kStbo * (t2 * t2);
  const float PAR = (0.92f * f.rsds) * 2.3f;
  d.tas = f.tas;
  d.forc_rain = M::div(1.0E3f * f.pr, kRhow);
  d.rain_dt = d.forc_rain * g.dt;
  const float lamb = (2503.0f - 2.386f * (tak - kTf)) * 1.0E3f;
  /* HYDROLOGY.f90:232-263 */
  const float tsv = tak * (1.0f + f.huss * kDeltx);
  const float rho = M::div(f.ps, kRgas * tsv);
  const float tc = tak - kTf;
  const float tc237 = tc + 237.3f;
  const float es = 0.6108f * M::exp(M::div(17.27f * tc, tc237));
  float desatdT = M::div(4098.0f * es, tc237 * tc237);
  desatdT = M::div(desatdT * 18.0f, kGasc * tak);
  const float esat = M::div(es * 18.0f, kGasc * tak);
  d.desatdT = desatdT;
  d.VDD = esat * (1.0f - M::div(f.rhs, 100.0f));
  d.gamma = M::div(kCp * f.ps, lamb * 0.622f) * M::div(18.0E-3f, kGasc * tak);
  d.dg = desatdT + d.gamma;
  d.rhocp = rho * kCp;
  d.rhow_lamb = kRhow * lamb;
  /* :283-295, the parts that do not depend on beta */
  d.canopy_on = (lai > 0.0f) && (PAR > 0.0f);
  d.rsc_num = M::div(1.0f, M::div(PAR, PAR + 300.0f)) * 400.0f;
  d.lai2 = 2.0f * lai;
  d.p28 = M::pow(2.8f, -M::div(80.0f * fmaxf(0.0f, d.VDD), rho));
  d.rsc_floor = M::div(1.0f, M::div(M::div(lai, 2.7f) * 0.9f, M::div(rho * 1.0E3f, 18.0f)));
  /* :302-318 */
  d.rac = (lai > 0.0f) ? M::div(25.0f, 2.0f * lai) : 1.0E6f;
  if (lai <= 4.0f) {
    d.raa = 0.25f * lai * 42.0f + 0.25f * (4.0f - lai) * 34.0f;
    d.ras = 0.25f * lai * 128.0f + 0.25f * (4.0f - lai) * 49.0f;
  } else {
    d.raa = 42.0f;
    d.ras = 128.0f;
  }
  /* :326-330 */
  d.litter1000 = 1000.0f * lai_litter;
  d.litter10 = 10.0f + d.litter1000;
  /* :335-358 */
  const float Rnets = Rnet * M::exp(-0.7f * lai);
  const float G = 0.2f * Rnets;
  const float Rnet_m_Rnets = Rnet - Rnets;
  d.A = desatdT * (Rnet - G);
  d.raa_rac = d.raa + d.rac;
  d.raa_ras = d.raa + d.ras;
  d.pmc_num = d.A + M::div(d.rhocp * d.VDD - desatdT * d.rac * (Rnets - G), d.raa_rac);
  d.pms_num = d.A + M::div(d.rhocp * d.VDD - desatdT * d.ras * Rnet_m_Rnets, d.raa_ras);
  d.Ra = d.dg * d.raa;
  d.dg_ras = d.dg * d.ras;
  d.dg_rac = d.dg * d.rac;
  d.lec_a = desatdT * Rnet_m_Rnets;
  d.les_a = desatdT * (Rnets - G);
}

/* jwt: index of the layer right above the water table, HYDROLOGY.f90:499-508 */
H9_HD int find_jwt(const Geo& g, float zwt) {
  int jwt = NL;
  H9_UNROLL
  for (int I = NL; I >= 1; --I)
    if (zwt <= g.zim[I]) jwt = I - 1;
  return jwt;
}

/* arr[idx] for a register array indexed 0..N-1, without dynamic indexing */
template <int N>
H9_HD float pick(const float (&a)[N], int idx) {
  H9_ASSERT(idx >= 0 && idx < N);
  float r = a[0];
  H9_UNROLL
  for (int k = 1; k < N; ++k)
    if (idx == k) r = a[k];
  return r;
}

/* specific yield of layer I at water-table depth zwtmm, HYDROLOGY.f90:963-965 */
template <class M>
H9_HD float specific_yield(float ths, float psi, float bsw, float zwtmm) {
  float s_y = ths * (1.0f - M::pow(1.0f + M::div(zwtmm, -psi), M::div(-1.0f, bsw)));
  return fmaxf(s_y, 0.02f);
}

/* ---- HYDROLOGY.f90:141-1283: one sub-step of one cell -------------------- */
template <class M>
H9_HD uint32_t hydrology_step(const Geo& g, const Params& p, const Day& d, State& s, StepOut& o) {
  uint32_t fault = 0;
  const float dt = g.dt;
  float theta[NL];

  /* :141-151 */
  float w0 = d.rain_dt + s.wa;
  H9_UNROLL
  for (int i = 0; i < NL; ++i) {
    w0 = w0 + s.h2o[i];
    theta[i] = M::div(s.h2o[i], g.dzw[i + 1]);
  }

  /* SurfaceRunoff :182-212 */
  const float qflx_top_soil = d.forc_rain;
  const float fsat = p.fmax * M::exp(-0.5f * kFff * s.zwt);
  float qflx_surf = fsat * qflx_top_soil;

  /* beta from the previous sub-step's smp :269-276 */
  float beta = 0.0f;
  H9_UNROLL
  for (int i = 0; i < NL; ++i) {
    float b = 1.0f - M::div(s.smp[i] - g.zc[i + 1], -150000.0f);
    b = fminf(1.0f, b);
    b = fmaxf(0.0f, b);
    beta = beta + s.rootr[i] * b;
  }

  /* rsc :283-295 */
  float rsc;
  if (d.canopy_on && (beta > 0.0f)) {
    rsc = M::div(d.rsc_num, d.lai2 * beta * d.p28);
  } else {
    rsc = 1.0E6f;
  }
  rsc = fmaxf(rsc, d.rsc_floor);

  /* rss :325-331 */
  float rss;
  if (theta[0] <= 0.15f) {
    rss = d.litter10 * M::exp(0.3563f * 100.0f * (0.15f - theta[0]));
  } else {
    rss = 10.0f + d.litter1000 * (1.0f - M::div(theta[0], p.theta_s[0]));
  }

  /* two-source Penman-Monteith :344-389 */
  const float PMc = M::div(d.pmc_num, d.desatdT + d.gamma * (1.0f + M::div(rsc, d.raa_rac)));
  const float PMs = M::div(d.pms_num, d.desatdT + d.gamma * (1.0f + M::div(rss, d.raa_ras)));
  const float Rs = d.dg_ras + d.gamma * rss;
  const float Rc = d.dg_rac + d.gamma * rsc;
  const float Cc = M::div(1.0f, 1.0f + M::div(Rc * d.Ra, Rs * (Rc + d.Ra)));
  const float Cs = M::div(1.0f, 1.0f + M::div(Rs * d.Ra, Rc * (Rs + d.Ra)));
  const float LE = Cc * PMc + Cs * PMs;
  const float VDD0 = d.VDD + M::div((d.A - d.dg * LE) * d.raa, d.rhocp);
  const float LEc = M::div(d.lec_a + M::div(d.rhocp * VDD0, d.rac),
                           d.desatdT + d.gamma * (1.0f + M::div(rsc, d.rac)));
  const float LEs = M::div(d.les_a + M::div(d.rhocp * VDD0, d.ras),
                           d.desatdT + d.gamma * (1.0f + M::div(rss, d.ras)));
  const float qflx_tran_veg_col = M::div(LEc * 1.0E3f, d.rhow_lamb);
  float qflx_evap_grnd = M::div(LEs * 1.0E3f, d.rhow_lamb);

  /* evaporation limit :396-400 */
  float evap_max1 =
      M::div(g.dz[1] * (theta[0] - kWatmin), dt) - qflx_tran_veg_col * s.rootr[0];
  evap_max1 = fmaxf(0.0f, evap_max1);
  qflx_evap_grnd = fminf(evap_max1, qflx_evap_grnd);

  /* Infiltration :426-478 (frac_h2osfc == 0, so the (1-frac) factors are 1*x) */
  float qflx_in_soil = 1.0f * (qflx_top_soil - qflx_surf);
  qflx_in_soil = qflx_in_soil - 1.0f * qflx_evap_grnd;
  const float qinmax = (1.0f - fsat) * fminf(fminf(p.hksat[0], p.hksat[1]), p.hksat[2]);
  const float qflx_infl_excess = fmaxf(0.0f, qflx_in_soil - 1.0f * qinmax);
  const float qflx_infl = qflx_in_soil - qflx_infl_excess;
  qflx_surf = qflx_surf + qflx_infl_excess;

  /* SoilWater :492-508 */
  float zwtmm = 1000.0f * s.zwt;
  int jwt = find_jwt(g, s.zwt);

  /* equilibrium profile :517-590; zq[0..7] layers 1..8, zq[8] the aquifer node */
  float zq[NL + 1];
  H9_UNROLL
  for (int i = 0; i < NL; ++i) {
    const float zlo = g.zi[i], zhi = g.zi[i + 1];
    const float ths = p.theta_s[i], psi = p.psi_s[i], bsw = p.bsw[i];
    const float npsi = -psi;
    float vol;
    if (zwtmm <= zlo) {
      vol = ths;
    } else {
      const float e1 = 1.0f - M::div(1.0f, bsw);
      const float temp0 = M::pow(M::div((npsi + zwtmm) - zlo, npsi), e1);
      if ((zwtmm < zhi) && (zwtmm > zlo)) { /* water table inside the layer :525-542 */
        const float voleq1 = M::div(M::div(psi * ths, e1), zwtmm - zlo) * (1.0f - temp0);
        vol = M::div(voleq1 * (zwtmm - zlo) + ths * (zhi - zwtmm), zhi - zlo);
        vol = fminf(ths, vol);
        vol = fmaxf(vol, 0.0f);
      } else { /* water table below the layer :548-558 */
        const float tempi = M::pow(M::div((npsi + zwtmm) - zhi, npsi), e1);
        vol = M::div(M::div(psi * ths, e1), zhi - zlo) * (tempi - temp0);
        vol = fmaxf(vol, 0.0f);
        vol = fminf(ths, vol);
      }
    }
    float z = psi * M::pow(fmaxf(M::div(vol, ths), 0.01f), -bsw);
    zq[i] = fmaxf(kSmpmin, z);
  }
  zq[NL] = 0.0f;
  if (jwt == NL) { /* :576-590 */
    const float ths = p.theta_s[NL - 1], psi = p.psi_s[NL - 1], bsw = p.bsw[NL - 1];
    const float npsi = -psi, zhi = g.zi[NL];
    const float e1 = 1.0f - M::div(1.0f, bsw);
    const float temp0 = M::pow(M::div((npsi + zwtmm) - zhi, npsi), e1);
    float vol = M::div(M::div(psi * ths, e1), zwtmm - zhi) * (1.0f - temp0);
    vol = fmaxf(vol, 0.0f);
    vol = fminf(ths, vol);
    float z = psi * M::pow(fmaxf(M::div(vol, ths), 0.01f), -bsw);
    zq[NL] = fmaxf(kSmpmin, z);
  }

  /* hk, dhkdw, smp, dsmpdw :598-639 */
  float hk[NL], dhkdw[NL], dsmpdw[NL];
  H9_UNROLL
  for (int i = 0; i < NL; ++i) {
    const int ip = (i + 1 < NL) ? i + 1 : NL - 1;
    float s1 = M::div(0.5f * (theta[i] + theta[ip]), 0.5f * (p.theta_s[i] + p.theta_s[ip]));
    s1 = fminf(1.0f, s1);
    const float s2 = p.hksat[i] * M::pow(s1, 2.0f * p.bsw[i] + 2.0f);
    hk[i] = s1 * s2;
    dhkdw[i] = (2.0f * p.bsw[i] + 3.0f) * s2 * M::div(1.0f, p.theta_s[i] + p.theta_s[ip]);
    float s_node = fmaxf(M::div(theta[i], p.theta_s[i]), 0.01f);
    s_node = fminf(1.0f, s_node);
    float sm = p.psi_s[i] * M::pow(s_node, -p.bsw[i]);
    sm = fmaxf(kSmpmin, sm);
    s.smp[i] = sm;
    dsmpdw[i] = M::div((-p.bsw[i]) * sm, s_node * p.theta_s[i]);
  }

  /* aquifer node :645-650 */
  const float zc9 = 0.5f * (zwtmm + g.zc[NL]);
  const float dz9 = (jwt < NL) ? g.dz[NL] : (zwtmm - g.zc[NL]);

  /* tridiagonal rows :661-799; index 0..8 = rows 1..9 */
  float amx[NL + 1], bmx[NL + 1], cmx[NL + 1], rmx[NL + 1];
  {
    const float den = g.zc[2] - g.zc[1];
    const float dzq = zq[1] - zq[0];
    const float num = (s.smp[1] - s.smp[0]) - dzq;
    const float qout = M::div(-hk[0] * num, den);
    const float dqodw1 = M::div(-(-hk[0] * dsmpdw[0] + num * dhkdw[0]), den);
    const float dqodw2 = M::div(-(hk[0] * dsmpdw[1] + num * dhkdw[0]), den);
    rmx[0] = qflx_infl - qout - qflx_tran_veg_col * s.rootr[0];
    amx[0] = 0.0f;
    bmx[0] = g.dzdt[1] + dqodw1;
    cmx[0] = dqodw2;
  }
  H9_UNROLL
  for (int i = 1; i < NL - 1; ++i) { /* rows 2..7 */
    float den = g.zc[i + 1] - g.zc[i];
    float dzq = zq[i] - zq[i - 1];
    float num = s.smp[i] - s.smp[i - 1] - dzq;
    const float qin = M::div(-hk[i - 1] * num, den);
    const float dqidw0 = M::div(-(-hk[i - 1] * dsmpdw[i - 1] + num * dhkdw[i - 1]), den);
    const float dqidw1 = M::div(-(hk[i - 1] * dsmpdw[i] + num * dhkdw[i - 1]), den);
    den = g.zc[i + 2] - g.zc[i + 1];
    dzq = zq[i + 1] - zq[i];
    num = (s.smp[i + 1] - s.smp[i]) - dzq;
    const float qout = M::div(-hk[i] * num, den);
    const float dqodw1 = M::div(-(-hk[i] * dsmpdw[i] + num * dhkdw[i]), den);
    const float dqodw2 = M::div(-(hk[i] * dsmpdw[i + 1] + num * dhkdw[i]), den);
    rmx[i] = qin - qout - qflx_tran_veg_col * s.rootr[i];
    amx[i] = -dqidw0;
    bmx[i] = g.dzdt[i + 1] - dqidw1 + dqodw1;
    cmx[i] = dqodw2;
  }
  {
    const int i = NL - 1; /* row 8 */
    float den = g.zc[NL] - g.zc[NL - 1];
    float dzq = zq[i] - zq[i - 1];
    float num = s.smp[i] - s.smp[i - 1] - dzq;
    const float qin = M::div(-hk[i - 1] * num, den);
    const float dqidw0 = M::div(-(-hk[i - 1] * dsmpdw[i - 1] + num * dhkdw[i - 1]), den);
    const float dqidw1 = M::div(-(hk[i - 1] * dsmpdw[i] + num * dhkdw[i - 1]), den);
    if (NL > jwt) { /* water table in the soil column :712-735 */
      rmx[i] = qin - 0.0f - qflx_tran_veg_col * s.rootr[i];
      amx[i] = -dqidw0;
      bmx[i] = g.dzdt[NL] - dqidw1 + 0.0f;
      cmx[i] = 0.0f;
      rmx[NL] = 0.0f;
      amx[NL] = 0.0f;
      bmx[NL] = M::div(dz9, dt);
      cmx[NL] = 0.0f;
    } else { /* water table below the soil column :737-799 */
      float s_node = fmaxf(0.5f * (1.0f + M::div(theta[i], p.theta_s[i])), 0.01f);
      s_node = fminf(1.0f, s_node);
      float smp1 = p.psi_s[i] * M::pow(s_node, -p.bsw[i]);
      smp1 = fmaxf(kSmpmin, smp1);
      const float dsmpdw1 = M::div(-p.bsw[i] * smp1, s_node * p.theta_s[i]);
      den = zc9 - g.zc[NL];
      dzq = zq[NL] - zq[i];
      num = smp1 - s.smp[i] - dzq;
      const float qout = M::div(-hk[i] * num, den);
      const float dqodw1 = M::div(-(-hk[i] * dsmpdw[i] + num * dhkdw[i]), den);
      const float dqodw2 = M::div(-(hk[i] * dsmpdw1 + num * dhkdw[i]), den);
      rmx[i] = qin - qout - qflx_tran_veg_col * s.rootr[i];
      amx[i] = -dqidw0;
      bmx[i] = g.dzdt[NL] - dqidw1 + dqodw1;
      cmx[i] = dqodw2;
      /* aquifer row: den/num unchanged, qin = qout :786-796 */
      const float dqidw0_9 = dqodw1;
      const float dqidw1_9 = dqodw2;
      rmx[NL] = qout - 0.0f;
      amx[NL] = -dqidw0_9;
      bmx[NL] = M::div(dz9, dt) - dqidw1_9 + 0.0f;
      cmx[NL] = 0.0f;
    }
  }

  /* Thomas :806-837 (reference STOPs on a zero pivot; here: flag) */
  float dwat2[NL + 1], gam[NL + 1];
  if (bmx[0] == 0.0f) fault |= FAULT_PIVOT1;
  float bet = bmx[0];
  dwat2[0] = M::div(rmx[0], bet);
  gam[0] = 0.0f;
  H9_UNROLL
  for (int i = 1; i <= NL; ++i) {
    gam[i] = M::div(cmx[i - 1], bet);
    bet = bmx[i] - amx[i] * gam[i];
    if (bet == 0.0f) fault |= FAULT_PIVOT2;
    dwat2[i] = M::div(rmx[i] - amx[i] * dwat2[i - 1], bet);
  }
  H9_UNROLL
  for (int i = NL - 1; i >= 0; --i) dwat2[i] = dwat2[i] - gam[i + 1] * dwat2[i + 1];

  /* :845-850 */
  H9_UNROLL
  for (int i = 0; i < NL; ++i) s.h2o[i] = s.h2o[i] + dwat2[i] * g.dz[i + 1];

  /* recharge :856-904 */
  float qcharge;
  if (jwt < NL) {
    /* layer jwt+1 (1-based) == index jwt; MAX(1,jwt) (1-based) == index max(jwt,1)-1 */
    const float th_j = pick<NL>(theta, jwt);
    const float ths_j = pick<NL>(p.theta_s, jwt);
    const float hks_j = pick<NL>(p.hksat, jwt);
    const float bsw_j = pick<NL>(p.bsw, jwt);
    const int jm = (jwt > 1 ? jwt : 1) - 1;
    const float s_node = fmaxf(M::div(th_j, ths_j), 0.01f);
    const float s1 = fminf(1.0f, s_node);
    const float ka = hks_j * M::pow(s1, 2.0f * bsw_j + 3.0f);
    const float smp1 = fmaxf(kSmpmin, pick<NL>(s.smp, jm));
    float zq_j = zq[0];
    H9_UNROLL
    for (int k = 1; k < NL; ++k)
      if (jm == k) zq_j = zq[k];
    const float wh = smp1 - zq_j;
    if (jwt == 0) {
      qcharge = M::div(-ka * (0.0f - wh), zwtmm + 1.0f);
    } else {
      float zc_j = g.zc[1];
      H9_UNROLL
      for (int k = 2; k <= NL; ++k)
        if (jwt == k) zc_j = g.zc[k];
      qcharge = M::div(-ka * (0.0f - wh), (zwtmm - zc_j) * 2.0f);
    }
    qcharge = fmaxf(g.q10_lo, qcharge);
    qcharge = fminf(g.q10_hi, qcharge);
  } else {
    qcharge = M::div(dwat2[NL] * dz9, dt);
  }

  /* Drainage :923-940 */
  jwt = find_jwt(g, s.zwt);
  float rous = specific_yield<M>(p.theta_s[NL - 1], p.psi_s[NL - 1], p.bsw[NL - 1], zwtmm);

  if (jwt == NL) { /* :946-951 */
    s.wa = s.wa + qcharge * dt;
    s.zwt = s.zwt - M::div(M::div(qcharge * dt, 1000.0f), rous);
  } else { /* :953-1009; zwtmm stays the stale value of :492 inside the loops */
    float qcharge_tot = qcharge * dt;
    if (qcharge_tot > 0.0f) { /* rising, layers jwt+1 .. 1 */
      bool done = false;
      H9_UNROLL
      for (int I = NL; I >= 1; --I) {
        if ((I <= jwt + 1) && !done) {
          const float s_y = specific_yield<M>(p.theta_s[I - 1], p.psi_s[I - 1], p.bsw[I - 1], zwtmm);
          float ql = fminf(qcharge_tot, s_y * (zwtmm - g.zi[I - 1]));
          ql = fmaxf(ql, 0.0f);
          if (s_y > 0.0f) s.zwt = s.zwt - M::div(M::div(ql, s_y), 1000.0f);
          qcharge_tot = qcharge_tot - ql;
          if (qcharge_tot <= 0.0f) done = true;
        }
      }
    } else { /* deepening, layers jwt+1 .. 8 */
      bool done = false;
      H9_UNROLL
      for (int I = 1; I <= NL; ++I) {
        if ((I >= jwt + 1) && !done) {
          const float s_y = specific_yield<M>(p.theta_s[I - 1], p.psi_s[I - 1], p.bsw[I - 1], zwtmm);
          float ql = fmaxf(qcharge_tot, -s_y * (g.zi[I] - zwtmm));
          ql = fminf(ql, 0.0f);
          qcharge_tot = qcharge_tot - ql;
          if (qcharge_tot >= 0.0f) {
            s.zwt = s.zwt - M::div(M::div(ql, s_y), 1000.0f);
            done = true;
          } else {
            s.zwt = g.zim[I];
          }
        }
      }
      if (qcharge_tot > 0.0f) s.zwt = s.zwt - M::div(M::div(qcharge_tot, 1000.0f), rous);
    }
    jwt = find_jwt(g, s.zwt); /* :1000-1007 */
  }

  zwtmm = 1000.0f * s.zwt; /* :1015 */

  /* baseflow :1024-1035 */
  float rsub_top = 5.5E-3f * M::exp(-kFff * s.zwt);
  rous = specific_yield<M>(p.theta_s[NL - 1], p.psi_s[NL - 1], p.bsw[NL - 1], zwtmm);

  if (jwt == NL) { /* :1048-1058; jwt is not recomputed on this path */
    s.wa = s.wa - rsub_top * dt;
    s.zwt = s.zwt + M::div(M::div(rsub_top * dt, 1000.0f), rous);
    s.h2o[NL - 1] = s.h2o[NL - 1] + fmaxf(0.0f, s.wa - 5000.0f);
    s.wa = fminf(s.wa, 5000.0f);
  } else { /* :1060-1118 */
    float rsub_top_tot = -rsub_top * dt;
    if (rsub_top_tot > 0.0f) {
      fault |= FAULT_RSUB; /* reference STOPs :1068-1071 */
    } else {
      bool done = false;
      H9_UNROLL
      for (int I = 1; I <= NL; ++I) {
        if ((I >= jwt + 1) && !done) {
          const float s_y = specific_yield<M>(p.theta_s[I - 1], p.psi_s[I - 1], p.bsw[I - 1], zwtmm);
          float rl = fmaxf(rsub_top_tot, -(s_y * (g.zi[I] - zwtmm)));
          rl = fminf(rl, 0.0f);
          s.h2o[I - 1] = s.h2o[I - 1] + rl;
          rsub_top_tot = rsub_top_tot - rl;
          if (rsub_top_tot >= 0.0f) {
            s.zwt = s.zwt - M::div(M::div(rl, s_y), 1000.0f);
            done = true;
          } else {
            s.zwt = g.zim[I];
          }
        }
      }
      /* residual, unconditional :1100-1101 */
      s.zwt = s.zwt - M::div(M::div(rsub_top_tot, 1000.0f), rous);
      s.wa = s.wa + rsub_top_tot;
    }
    jwt = find_jwt(g, s.zwt); /* :1110-1116 */
  }

  s.zwt = fmaxf(0.0f, s.zwt);  /* :1122 */
  s.zwt = fminf(80.0f, s.zwt); /* :1123 */

  /* excess cascade :1131-1152 (eff_porosity = MAX(0.01,theta_s) :430) */
  H9_UNROLL
  for (int i = NL - 1; i >= 1; --i) {
    const float cap = fmaxf(0.01f, p.theta_s[i]) * g.dz[i + 1];
    const float xsi = fmaxf(s.h2o[i] - cap, 0.0f);
    s.h2o[i] = fminf(cap, s.h2o[i]);
    s.h2o[i - 1] = s.h2o[i - 1] + xsi;
  }
  const float cap1 = fmaxf(0.0f, p.theta_s[0] * g.dz[1]);
  const float xs1 = fmaxf(fmaxf(s.h2o[0], 0.0f) - cap1, 0.0f);
  s.h2o[0] = fminf(cap1, s.h2o[0]);
  const float qflx_rsub_sat = M::div(xs1, dt);

  /* dryness repair :1161-1205 */
  float xs;
  H9_UNROLL
  for (int i = 0; i < NL - 1; ++i) {
    if (s.h2o[i] < kWatmin) {
      xs = kWatmin - s.h2o[i];
      if (i + 1 == jwt) s.zwt = s.zwt + M::div(M::div(xs, fmaxf(0.01f, p.theta_s[i])), 1000.0f);
    } else {
      xs = 0.0f;
    }
    s.h2o[i] = s.h2o[i] + xs;
    s.h2o[i + 1] = s.h2o[i + 1] - xs;
  }
  if (s.h2o[NL - 1] < kWatmin) {
    xs = kWatmin - s.h2o[NL - 1];
    bool done = false;
    H9_UNROLL
    for (int j = NL - 2; j >= 0; --j) {
      if (!done) {
        const float avail = fmaxf(s.h2o[j] - kWatmin - xs, 0.0f);
        if (avail >= xs) {
          s.h2o[NL - 1] = s.h2o[NL - 1] + xs;
          s.h2o[j] = s.h2o[j] - xs;
          xs = 0.0f;
          done = true;
        } else {
          s.h2o[NL - 1] = s.h2o[NL - 1] + avail;
          s.h2o[j] = s.h2o[j] - avail;
          xs = xs - avail;
        }
      }
    }
  } else {
    xs = 0.0f;
  }
  s.h2o[NL - 1] = s.h2o[NL - 1] + xs;   /* :1205 */
  rsub_top = rsub_top - M::div(xs, dt); /* :1211 */

  /* balance :1221-1244 */
  float w1 = (1.0f * (qflx_surf + qflx_evap_grnd + qflx_tran_veg_col) + rsub_top + qflx_rsub_sat) * dt +
             s.wa;
  H9_UNROLL
  for (int i = 0; i < NL; ++i) w1 = w1 + s.h2o[i];
  const float imb = w1 - w0;
  if (!(fabsf(imb) <= 0.1f)) fault |= FAULT_IMBAL;

  /* :1282-1283 */
  const float r1 = qflx_surf * dt, r2 = rsub_top * dt;
  s.rnf_sum = s.rnf_sum + r1;
  s.rnf_sum = s.rnf_sum + r2;

  o.qflx_tran_veg_col = qflx_tran_veg_col;
  o.qflx_evap_grnd = qflx_evap_grnd;
  o.rnf_inc = r1 + r2;
  o.imbalance = imb;
  o.jwt = jwt;
  return fault;
}

/* theta as HYDROLOGY leaves it at the end of a call, :1233-1234 */
template <class M>
H9_HD float theta_diag(const Geo& g, float h2o, int i) {
  return M::div(fmaxf(h2o, 1.0E-6f), g.dzw[i + 1]);
}

/* root profile of one plant added to rootr, INIT.f90:791-797 == GROW.f90:176-182 */
template <class M>
H9_HD void root_profile(const Geo& g, float rdepth, float (&rootr)[NL]) {
  const float decay = M::exp(M::div(M::log(0.1f), M::div(rdepth, 10.0f)));
  float prev = M::pow(decay, g.zi10[0]);
  H9_UNROLL
  for (int i = 0; i < NL; ++i) {
    const float cur = M::pow(decay, g.zi10[i + 1]);
    rootr[i] = rootr[i] + (1.0f - cur) - (1.0f - prev);
    prev = cur;
  }
}

struct GrowOut {
  float npp, w_i, fT;
};

/* ---- GROW.f90:55-201: one day of one cell -------------------------------- */
template <class M>
H9_HD void grow_day(const Geo& g, float tas, State& s, GrowOut& o) {
  float w_i = 0.0f;
  H9_UNROLL
  for (int i = 0; i < NL; ++i) { /* :55-62 */
    float w = M::div(-150000.0f - s.smp[i], -150000.0f - (-50000.0f));
    w = fmaxf(0.0f, w);
    w = fminf(1.0f, w);
    w_i = w_i + s.rootr[i] * w;
  }
  float fT; /* :66-72 */
  if ((tas - kTf) > 18.0f) {
    const float q = M::div(fabsf(tas - kTf - 18.0f), 21.0f);
    fT = 1.0f - q * q;
  } else {
    const float q = M::div(fabsf(tas - kTf - 18.0f), 25.0f);
    fT = 1.0f - q * q;
    fT = fmaxf(0.0f, fT);
    fT = fminf(1.0f, fT);
  }
  H9_UNROLL
  for (int i = 0; i < NL; ++i) s.rootr[i] = 0.0f; /* :76 */
  float npp = 0.0f;
  /* nplants_max == 1 (SHARED.f90:63): the plant loop :82-188 runs 0 or 1 times */
  if (s.nplants >= 1) {
    const float grow_plant_mass = (1000.0f / 365.0f) * w_i * fT; /* :90 */
    const float grow_foliage_mass = M::div(grow_plant_mass, 3.3f);
    const float loss_plant_mass = (0.1f / 365.0f) * s.plant_mass; /* :134 */
    float loss_foliage_mass =
        M::div((1.0f / 365.0f) * s.plant_foliage_mass, fminf(1.0f, fmaxf(0.01f, w_i))); /* :136 */
    if (w_i < 0.6f) loss_foliage_mass = 0.1f * s.plant_foliage_mass;                    /* :138 */
    const float dplant_mass = grow_plant_mass - loss_plant_mass;
    const float dplant_foliage_mass = grow_foliage_mass - loss_foliage_mass;
    s.plant_mass = s.plant_mass + dplant_mass;
    s.plant_foliage_mass = s.plant_foliage_mass + dplant_foliage_mass;
    s.plant_length = M::pow(M::div(400.0f * s.plant_mass, 3.142E-3f), 1.0f / 3.0f); /* :155 */
    const float dLAI = dplant_foliage_mass * kSla;                                  /* :161 */
    s.lai = s.lai + dLAI;
    s.lai = fmaxf(0.001f, s.lai);
    s.lai_litter = s.lai_litter + fmaxf(0.0f, dLAI); /* :167 */
    s.rdepth = 0.3f * s.plant_length;                /* :171 */
    root_profile<M>(g, s.rdepth, s.rootr);           /* :176-182 */
    npp = npp + dplant_mass;                         /* :186 */
  }
  s.lai_litter = s.lai_litter - 0.02f * s.lai_litter; /* :201 */
  o.npp = npp;
  o.w_i = w_i;
  o.fT = fT;
}

} /* namespace h9 */
#endif
