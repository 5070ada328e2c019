// mwrt_common.cuh -- shared device helpers for libmwrt (sm_100a only).
//
// Everything here is fp64: the path is bound by the FP64 pipe (64 lanes/clk/SM on B200) and MUFU,
// not by HBM or tensor cores (DESIGN.md "Roofline").  The helpers trade IEEE corner cases the
// physics cannot reach (denormals, infinities in denominators) for fewer FP64 issue slots.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "../../include/mwrt.h"

namespace mwrt {

// physical constants -- the values PyRTlib's utils.constants() carries (oracle/pyrtlib_scalar.py)
constexpr double kPlanckH = 6.6260755e-34;
constexpr double kBoltzmannK = 1.380658e-23;
constexpr double kLightC = 299792458.0;
constexpr double kTCosmic = 2.73;
constexpr double kEarthRadius = 6370.949;
constexpr double kRWatVap = 461.5;
constexpr double kSqrtPi = 1.77245385090551603;
constexpr double kExpMax = 125.0;
constexpr double kDb2Np = 0.2302585092994046;  // ln(10)/10
constexpr double kRvapAbs = (0.01 * 8.314510) / 18.01528;

// Device-side view of a model: scalars by value, line tables by device pointer (column-major,
// MWRT_H2O_NCOL x n_h2o then MWRT_O2_NCOL x n_o2 in ONE allocation so a CTA stages it with one loop).
struct ModelDev {
  int n_h2o, n_o2;
  int h2o_sd, o2_mixing, o2_dens118, o2_sd118, n2_fdep, liq_model;
  double h2o_reftline, h2o_reftcon, h2o_cf, h2o_xcf, h2o_cs, h2o_xcs, h2o_pvap_div, h2o_den_const,
      h2o_line_const;
  double o2_x, o2_wb300, o2_pvap_div, o2_h2o_broad, o2_nonres, o2_abs_const, o2_iso;
  double n2_l, n2_m, n2_n;
  const double* tables;  // device
  int n_table;           // doubles in `tables`
};

// ---- reciprocal / division without the IEEE slow path --------------------------------------
// MUFU.RCP64H seed (~20 bits) + one cubic Newton step (3 DFMA) -> ~2^-60 before rounding.
// Valid for normal, finite x (all line-shape denominators are >= width^2 > 0).
__device__ __forceinline__ double rcp_fast(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  double t = fma(e, e, e);
  return fma(r, t, r);
}
// a/b with one residual correction (result within 1 ulp of IEEE)
__device__ __forceinline__ double div_fast(double a, double b) {
  double r = rcp_fast(b);
  double q = a * r;
  double rem = fma(-b, q, a);
  return fma(rem, r, q);
}

// ---- branch-free exp ---------------------------------------------------------------------------
// exp(x) for |x| < 700 (callers guarantee the range; no NaN/inf/denormal handling).  Cody-Waite
// reduction by ln2 (k via the 1.5*2^52 rounding constant), degree-13 Taylor on |r| <= ln2/2
// (truncation 4e-18), scaling by an integer add into the exponent field.  17 FP64-pipe instructions
// and no branch, so the scheduler can interleave it with neighbouring line-shape arithmetic.
__device__ __forceinline__ double exp_nb(double x) {
  const double kMagic = 6755399441055744.0;  // 1.5 * 2^52
  double t = fma(x, 1.4426950408889634, kMagic);
  const int k = __double2loint(t);
  t -= kMagic;
  double r = fma(-t, 6.93147180369123816490e-01, x);
  r = fma(-t, 1.90821492927058770002e-10, r);
  double p = 1.6059043836821613e-10;           // 1/13!
  p = fma(p, r, 2.08767569878681e-09);         // 1/12!
  p = fma(p, r, 2.505210838544172e-08);        // 1/11!
  p = fma(p, r, 2.755731922398589e-07);        // 1/10!
  p = fma(p, r, 2.7557319223985893e-06);       // 1/9!
  p = fma(p, r, 2.48015873015873e-05);         // 1/8!
  p = fma(p, r, 1.984126984126984e-04);        // 1/7!
  p = fma(p, r, 1.388888888888889e-03);        // 1/6!
  p = fma(p, r, 8.333333333333333e-03);        // 1/5!
  p = fma(p, r, 4.1666666666666664e-02);       // 1/4!
  p = fma(p, r, 1.6666666666666666e-01);       // 1/3!
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}

// ---- branch-free natural log ---------------------------------------------------------------------
// log(x) for normal positive x (callers substitute 1.0 for anything else and select afterwards).
// x = 2^e * m with m in [sqrt(1/2), sqrt(2)); log m = 2 atanh(s), s = (m-1)/(m+1), |s| <= 0.1716,
// odd series to s^21 (truncation 2e-17).  ~22 FP64-pipe instructions + 1 MUFU, no branch.
__device__ __forceinline__ double log_nb(double x) {
  int hi = __double2hiint(x);
  const int lo = __double2loint(x);
  int e = (hi >> 20) - 1023;
  hi = (hi & 0x000fffff) | 0x3ff00000;            // m in [1, 2)
  const bool big = hi >= 0x3ff6a09f;              // m >= sqrt(2) (top mantissa bits)
  hi = big ? hi - 0x00100000 : hi;                // m/2
  e = big ? e + 1 : e;
  const double m = __hiloint2double(hi, lo);
  const double s = (m - 1.0) * rcp_fast(m + 1.0);
  const double z = s * s;
  double q = 1.0 / 21.0;
  q = fma(q, z, 1.0 / 19.0);
  q = fma(q, z, 1.0 / 17.0);
  q = fma(q, z, 1.0 / 15.0);
  q = fma(q, z, 1.0 / 13.0);
  q = fma(q, z, 1.0 / 11.0);
  q = fma(q, z, 1.0 / 9.0);
  q = fma(q, z, 1.0 / 7.0);
  q = fma(q, z, 1.0 / 5.0);
  q = fma(q, z, 1.0 / 3.0);
  const double s2 = s + s;
  const double ed = (double)e;
  // e*ln2_hi + (2s + (2s*z*q + e*ln2_lo)); ln2_hi has its low 21 bits clear so e*ln2_hi is exact
  const double tail = fma(s2 * z, q, ed * 1.90821492927058770002e-10);
  return fma(ed, 6.93147180369123816490e-01, s2 + tail);
}

// The same exp over N independent arguments, written stage by stage so that the N dependency chains
// are interleaved in program order (the FP64 pipe needs ~4 independent DFMAs in flight per scheduler;
// left to itself the compiler serialises the chains to save registers).
// Ordered FP64 ops: `asm volatile` keeps the compiler front end from re-serialising the N chains
// (it otherwise schedules depth-first to save registers; measured: 52% of FP64 instructions issued
// back-to-back dependent).  ptxas still owns the final schedule but starts from this order.
#define MWRT_FMA(d, a, b, c) asm volatile("fma.rn.f64 %0, %1, %2, %3;" : "=d"(d) : "d"(a), "d"(b), "d"(c))
#define MWRT_MUL(d, a, b) asm volatile("mul.rn.f64 %0, %1, %2;" : "=d"(d) : "d"(a), "d"(b))
#define MWRT_ADD(d, a, b) asm volatile("add.rn.f64 %0, %1, %2;" : "=d"(d) : "d"(a), "d"(b))

template <int N>
__device__ __forceinline__ void exp_nb_vec(const double (&x)[N], double (&y)[N]) {
  const double kMagic = 6755399441055744.0;
  double t[N], r[N], p[N];
  int k[N];
#pragma unroll
  for (int a = 0; a < N; ++a) MWRT_FMA(t[a], x[a], 1.4426950408889634, kMagic);
#pragma unroll
  for (int a = 0; a < N; ++a) {
    k[a] = __double2loint(t[a]);
    MWRT_ADD(t[a], t[a], -kMagic);
  }
#pragma unroll
  for (int a = 0; a < N; ++a) MWRT_FMA(r[a], t[a], -6.93147180369123816490e-01, x[a]);
#pragma unroll
  for (int a = 0; a < N; ++a) MWRT_FMA(r[a], t[a], -1.90821492927058770002e-10, r[a]);
#pragma unroll
  for (int a = 0; a < N; ++a) MWRT_FMA(p[a], r[a], 1.6059043836821613e-10, 2.08767569878681e-09);
  const double c[11] = {2.505210838544172e-08, 2.755731922398589e-07, 2.7557319223985893e-06,
                        2.48015873015873e-05,  1.984126984126984e-04, 1.388888888888889e-03,
                        8.333333333333333e-03, 4.1666666666666664e-02, 1.6666666666666666e-01,
                        0.5,                   1.0};
#pragma unroll
  for (int i = 0; i < 11; ++i) {
#pragma unroll
    for (int a = 0; a < N; ++a) MWRT_FMA(p[a], p[a], r[a], c[i]);
  }
#pragma unroll
  for (int a = 0; a < N; ++a) MWRT_FMA(p[a], p[a], r[a], 1.0);
#pragma unroll
  for (int a = 0; a < N; ++a)
    y[a] = __hiloint2double(__double2hiint(p[a]) + (k[a] << 20), __double2loint(p[a]));
}

// N reciprocals, staged like exp_nb_vec
template <int N>
__device__ __forceinline__ void rcp_fast_vec(const double (&x)[N], double (&y)[N]) {
  double r[N], e[N];
#pragma unroll
  for (int a = 0; a < N; ++a) asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r[a]) : "d"(x[a]));
#pragma unroll
  for (int a = 0; a < N; ++a) MWRT_FMA(e[a], -x[a], r[a], 1.0);
#pragma unroll
  for (int a = 0; a < N; ++a) MWRT_FMA(e[a], e[a], e[a], e[a]);
#pragma unroll
  for (int a = 0; a < N; ++a) MWRT_FMA(y[a], r[a], e[a], r[a]);
}

// ---- minimal complex arithmetic (speed-dependent line shape, liquid permittivity) -----------
struct cplx {
  double re, im;
};
__device__ __forceinline__ cplx cmake(double a, double b) { return cplx{a, b}; }
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return cplx{a.re + b.re, a.im + b.im}; }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return cplx{a.re - b.re, a.im - b.im}; }
__device__ __forceinline__ cplx cmul(cplx a, cplx b) {
  return cplx{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
}
__device__ __forceinline__ cplx cscale(cplx a, double s) { return cplx{a.re * s, a.im * s}; }
__device__ __forceinline__ cplx cconj(cplx a) { return cplx{a.re, -a.im}; }
__device__ __forceinline__ cplx cdiv(cplx a, cplx b) {
  // Smith's algorithm (what C99/NumPy use) to stay close to the oracle's rounding
  if (fabs(b.re) >= fabs(b.im)) {
    double r = b.im / b.re;
    double d = 1.0 / (b.re + b.im * r);
    return cplx{(a.re + a.im * r) * d, (a.im - a.re * r) * d};
  } else {
    double r = b.re / b.im;
    double d = 1.0 / (b.re * r + b.im);
    return cplx{(a.re * r + a.im) * d, (a.im * r - a.re) * d};
  }
}
__device__ __forceinline__ cplx csqrt_(cplx z) {
  double m = hypot(z.re, z.im);
  if (m == 0.0) return cplx{0.0, 0.0};
  double t = sqrt(0.5 * (m + fabs(z.re)));
  if (z.re >= 0.0) return cplx{t, z.im / (2.0 * t)};
  return cplx{fabs(z.im) / (2.0 * t), copysign(t, z.im)};
}
__device__ __forceinline__ cplx cexp_(cplx z) {
  double e = exp(z.re), s, c;
  sincos(z.im, &s, &c);
  return cplx{e * c, e * s};
}
__device__ __forceinline__ cplx clog_(cplx z) { return cplx{log(hypot(z.re, z.im)), atan2(z.im, z.re)}; }

// Complex error function w(x+iy): Hui, Armstrong & Wray 6th-order rational approximation
// (Rosenkranz `dcerror`, as in pyrtlib.utils.dcerror).
__device__ __forceinline__ cplx dcerror(double x, double y) {
  const double a0 = 122.607931777104326, a1 = 214.382388694706425, a2 = 181.928533092181549,
               a3 = 93.155580458138441, a4 = 30.180142196210589, a5 = 5.912626209773153,
               a6 = 0.564189583562615;
  const double b0 = 122.607931773875350, b1 = 352.730625110963558, b2 = 457.334478783897737,
               b3 = 348.703917719495792, b4 = 170.354001821091472, b5 = 53.992906912940207,
               b6 = 10.479857114260399;
  cplx zh = cmake(fabs(y), -x);
  cplx as = cmake(a6 * zh.re + a5, a6 * zh.im);
  as = cmul(as, zh); as.re += a4;
  as = cmul(as, zh); as.re += a3;
  as = cmul(as, zh); as.re += a2;
  as = cmul(as, zh); as.re += a1;
  as = cmul(as, zh); as.re += a0;
  cplx bs = cmake(zh.re + b6, zh.im);
  bs = cmul(bs, zh); bs.re += b5;
  bs = cmul(bs, zh); bs.re += b4;
  bs = cmul(bs, zh); bs.re += b3;
  bs = cmul(bs, zh); bs.re += b2;
  bs = cmul(bs, zh); bs.re += b1;
  bs = cmul(bs, zh); bs.re += b0;
  cplx w = cdiv(as, bs);
  if (y >= 0.0) return w;
  cplx z = cmake(x, y);
  cplx z2 = cmul(z, z);
  cplx ex = cexp_(cmake(-z2.re, -z2.im));
  return cmake(2.0 * ex.re - w.re, 2.0 * ex.im + w.im);
}

// Real part of the speed-dependent resonant shape factor (H2O 22/183 GHz):
//   xc = (w0 - 1.5 w2 + i (df + 1.5 d2)) / (w2 - i d2);  sd = 2 (1 - sqrt(pi) xrt w(i xrt)) / (w2 - i d2)
__device__ __forceinline__ cplx sd_shape(double width0, double width2, double delta2, double df) {
  cplx cden = cmake(width2, -delta2);
  cplx xc = cdiv(cmake(width0 - 1.5 * width2, df + 1.5 * delta2), cden);
  cplx xrt = csqrt_(xc);
  cplx w = dcerror(-xrt.im, xrt.re);
  cplx pxw = cscale(cmul(xrt, w), kSqrtPi);
  cplx num = cmake(2.0 * (1.0 - pxw.re), -2.0 * pxw.im);
  return cdiv(num, cden);
}

// Goff-Gratch saturation vapour pressure over water [mb] (RTEquation.vapor)
__device__ __forceinline__ double goff_gratch_es(double tk) {
  double y = 373.16 / tk;
  double es = -7.90298 * (y - 1.0) + 5.02808 * log10(y) -
              1.3816e-07 * (exp10(11.344 * (1.0 - (1.0 / y))) - 1.0) +
              0.0081328 * (exp10(-3.49149 * (y - 1.0)) - 1.0) + 3.0057148979490314;  // log10(1013.246)
  return exp10(es);
}

}  // namespace mwrt

#define MWRT_CUDA_OK(call)                         \
  do {                                             \
    cudaError_t e__ = (call);                      \
    if (e__ != cudaSuccess) return mwrt::fail_cuda(e__, #call); \
  } while (0)

namespace mwrt {
int fail_cuda(cudaError_t e, const char* what);
int fail_arg(int code, const char* what);
}  // namespace mwrt
