// mwrt_absorption.cu -- kernel 1: gas (and cloud) absorption for every profile x frequency x level.
//
// Replaces, for a whole batch at once, what PyRTlib does inside TbCloudRTE.execute() per
// (angle, frequency, level): RTEquation.clearsky_absorption -> H2OAbsModel.h2o_absorption,
// O2AbsModel.o2_absorption, N2AbsModel.n2_absorption (SURVEY.md 8(a) rows 8,10; Appendix A.2-A.6).
// Cloud liquid/ice (row 13, A.7) is NOT here: on the fused path kernel 3 evaluates it inline for the
// cloudy levels (mwrt_cloud.cuh); the stage entry point uses the elementwise k_cloud_abs below.  The reference
// reaches it through `rte.execute()` at /root/reference/python_src/proc/PyRTlib_processing.py:126.
//
// Mapping (B200-first, not PyRTlib's loop order):
//   thread  = one level-point (profile p, level l); consecutive threads = consecutive levels, so the
//             p/t/e loads and the awet/adry stores of a warp are contiguous
//   grid.y  = block of FB frequencies; the thread keeps FB running line sums in registers
//   lines   = outer loop.  Everything that does not depend on frequency (widths, shifts, strengths,
//             mixing coefficients: ~130 exp() per level-point) is computed once per line and reused
//             for the FB frequencies, so the inner loop is ~11-13 DFMA-class ops + 1 MUFU.RCP64H per
//             (line, frequency): both Lorentz wings share ONE reciprocal.
//   tables  = the model's line tables are staged into shared memory once per CTA and read as
//             warp-uniform broadcasts; everything that depends on the frequency BLOCK only -- (f -/+ f_line)
//             in the oracle's rounding order, line classes and processing orders, the far-line grid --
//             is built once per plan by k_abs_blocks and copied by the CTAs.
//   far     = wide blocks (28/32 frequencies): lines far from the block are summed at 6 Chebyshev nodes
//             of the block and interpolated (see kNQ).
// Absorption does not depend on the viewing angle: it is computed once and reused for all angles
// (PyRTlib recomputes it per angle).
#include "mwrt_common.cuh"
#include "mwrt_cloud.cuh"
#include <stdlib.h>
#include <mutex>

namespace mwrt {

#define exp_tb_(x) exp_tb((x), etab)   // etab: the kernel's shared-memory exp table

#ifndef MWRT_ABS_THREADS
#define MWRT_ABS_THREADS 128   // measured: 64 and 256 threads per CTA (same warps per SM) are not faster
#endif
constexpr int kAbsThreads = MWRT_ABS_THREADS;
constexpr int kAllIn = 0, kAllOut = 1, kWing1Out = 2, kMixed = 3, kFarIn = 4;

// rare-path helpers are kept out of line: the hot loops must stay small enough for the
// instruction caches (the first version inlined them 14x and stalled on instruction fetch)
// Speed-dependent shapes with division-free complex arithmetic (reciprocals via MUFU + Newton, no
// IEEE slow-path calls, no branches): the 22 GHz line takes this path for every K-band channel in the
// lower troposphere, so it is not that rare (profiles/r01d: ~20 % of the kernel before this rewrite).
__device__ __forceinline__ cplx cdiv_f(cplx a, cplx b) {
  const double r = rcp_fast(fma(b.re, b.re, b.im * b.im));
  return cplx{fma(a.re, b.re, a.im * b.im) * r, fma(a.im, b.re, -a.re * b.im) * r};
}
__device__ __forceinline__ float rcp_f32(float x) {   // MUFU.RCP, ~1 ulp (fp32 line-shape mode)
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ cplx csqrt_f(cplx z) {  // principal root, z != 0
  const double m = sqrt_f(fma(z.re, z.re, z.im * z.im));
  const double t = sqrt_f(0.5 * (m + fabs(z.re)));
  const double o = 0.5 * z.im * rcp_fast(t);
  return (z.re >= 0.0) ? cplx{t, o} : cplx{fabs(o), copysign(t, z.im)};
}
// w(x + i y) for y >= 0 (always the case below: y = Re sqrt(.) >= 0)
__device__ __forceinline__ cplx dcerror_f(double x, double y) {
  const double a0 = 122.607931777104326, a1 = 214.382388694706425, a2 = 181.928533092181549,
               a3 = 93.155580458138441, a4 = 30.180142196210589, a5 = 5.912626209773153,
               a6 = 0.564189583562615;
  const double b0 = 122.607931773875350, b1 = 352.730625110963558, b2 = 457.334478783897737,
               b3 = 348.703917719495792, b4 = 170.354001821091472, b5 = 53.992906912940207,
               b6 = 10.479857114260399;
  const cplx zh = cmake(y, -x);
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
  return cdiv_f(as, bs);
}
// Speed-dependent shape, restructured for the hot correction loop (same function as sd_shape_re):
//   * 1/(w2 - i d2) and the frequency-independent part of xc are hoisted per line (SdLine);
//   * w(z) = P6(z)/Q7(z) has REAL coefficients, so both polynomials are evaluated at the complex z
//     with the quadratic-factor recurrence b_k = a_k + 2 Re(z) b_{k+1} - |z|^2 b_{k+2}
//     (P(z) = b_1 z + b_0 - 2 Re(z) b_1): 2 DFMA per coefficient instead of a complex multiply;
//     |z|^2 = |xc| falls out of the complex square root; sqrt(pi) is folded into P's coefficients;
//   * the square roots return 1/(2 sqrt) as a by-product, which replaces a reciprocal.
// ~74 FP64-pipe instructions per frequency instead of ~120.
struct SdLine {
  double a_re, a_im, icr, ici, icr2, ici2;
};
__device__ __forceinline__ SdLine sd_prepare(double width0, double width2, double delta2) {
  SdLine c;
  const double q = rcp_fast(fma(width2, width2, delta2 * delta2));
  c.icr = width2 * q;           // 1/(w2 - i d2) = (w2 + i d2)/(w2^2 + d2^2)
  c.ici = delta2 * q;
  const double A = width0 - 1.5 * width2, B0 = 1.5 * delta2;
  c.a_re = fma(A, c.icr, -B0 * c.ici);
  c.a_im = fma(A, c.ici, B0 * c.icr);
  c.icr2 = 2.0 * c.icr;
  c.ici2 = 2.0 * c.ici;
  return c;
}
// sqrt(x) and 1/(2 sqrt(x)) for normal x > 0
__device__ __forceinline__ void sqrt_pair(double x, double& s, double& h) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  h = 0.5 * y;
  s = x * y;
  double e = fma(-s, h, 0.5);
  s = fma(s, e, s);
  h = fma(h, e, h);
  e = fma(-s, h, 0.5);
  s = fma(s, e, s);
  h = fma(h, e, h);
  s = fma(fma(-s, s, x), h, s);
}
__device__ __forceinline__ double sd_shape_fast(const SdLine& c, double df) {
  // xc = (w0 - 1.5 w2 + i (df + 1.5 d2)) / (w2 - i d2)
  const double xr = fma(-df, c.ici, c.a_re), xi = fma(df, c.icr, c.a_im);
  // principal square root z = (zr, zi), |z|^2 = m
  double m, hm, t, ht;
  sqrt_pair(fma(xr, xr, xi * xi), m, hm);
  sqrt_pair(fma(0.5, m, 0.5 * fabs(xr)), t, ht);
  const double o = xi * ht;                       // 0.5 xi / t
  const bool pos = xr >= 0.0;
  const double zr = pos ? t : fabs(o);
  const double zi = pos ? o : copysign(t, xi);
  const double r = zr + zr;
  // sqrt(pi) * P6 (Hui, Armstrong & Wray numerator) and monic Q7
  constexpr double sp = kSqrtPi;
  constexpr double a0 = sp * 122.607931777104326, a1 = sp * 214.382388694706425,
                   a2 = sp * 181.928533092181549, a3 = sp * 93.155580458138441,
                   a4 = sp * 30.180142196210589, a5 = sp * 5.912626209773153,
                   a6 = sp * 0.564189583562615;
  constexpr double b0 = 122.607931773875350, b1 = 352.730625110963558, b2 = 457.334478783897737,
                   b3 = 348.703917719495792, b4 = 170.354001821091472, b5 = 53.992906912940207,
                   b6 = 10.479857114260399;
  double p2 = a6, p1 = fma(r, a6, a5), p0;
  p0 = fma(-m, p2, fma(r, p1, a4)); p2 = p1; p1 = p0;
  p0 = fma(-m, p2, fma(r, p1, a3)); p2 = p1; p1 = p0;
  p0 = fma(-m, p2, fma(r, p1, a2)); p2 = p1; p1 = p0;
  p0 = fma(-m, p2, fma(r, p1, a1)); p2 = p1; p1 = p0;
  p0 = fma(-m, p2, fma(r, p1, a0));               // p0 = b_0, p1 = b_1
  const double pre = fma(-p1, zr, p0), pim = p1 * zi;   // P(z) = b_1 z + b_0 - r b_1
  double q2 = 1.0, q1 = r + b6, q0;
  q0 = fma(r, q1, b5) - m; q2 = q1; q1 = q0;
  q0 = fma(-m, q2, fma(r, q1, b4)); q2 = q1; q1 = q0;
  q0 = fma(-m, q2, fma(r, q1, b3)); q2 = q1; q1 = q0;
  q0 = fma(-m, q2, fma(r, q1, b2)); q2 = q1; q1 = q0;
  q0 = fma(-m, q2, fma(r, q1, b1)); q2 = q1; q1 = q0;
  q0 = fma(-m, q2, fma(r, q1, b0));
  const double qre = fma(-q1, zr, q0), qim = q1 * zi;
  // w' = sqrt(pi) w(z) = P/Q;  pxw = z * w'
  const double rd = rcp_fast(fma(qre, qre, qim * qim));
  const double wre = fma(pre, qre, pim * qim) * rd, wim = fma(pim, qre, -pre * qim) * rd;
  const double pxr = fma(zr, wre, -zi * wim), pxi = fma(zr, wim, zi * wre);
  // Re[ 2 (1 - pxw) / (w2 - i d2) ]
  return fma(pxi, c.ici2, (1.0 - pxr) * c.icr2);
}
__device__ __noinline__ double sd_shape_re(double width0, double width2, double delta2, double df) {
  const cplx cden = cmake(width2, -delta2);
  const cplx xrt = csqrt_f(cdiv_f(cmake(width0 - 1.5 * width2, df + 1.5 * delta2), cden));
  const cplx pxw = cscale(cmul(xrt, dcerror_f(-xrt.im, xrt.re)), kSqrtPi);
  return cdiv_f(cmake(2.0 * (1.0 - pxw.re), -2.0 * pxw.im), cden).re;
}
__device__ __noinline__ double o2_sd118_sf1(double df, double d1, double y) {
  const double width2 = 0.076 * df;
  const double rw2 = rcp_fast(width2);
  const cplx xrt = csqrt_f(cmake((df - 1.5 * width2) * rw2, d1 * rw2));
  const cplx pxw = cscale(cmul(xrt, dcerror_f(-xrt.im, xrt.re)), kSqrtPi);
  const cplx one_m = cmake(2.0 * (1.0 - pxw.re), -2.0 * pxw.im);
  return cmul(cmake(1.0, y), one_m).re * rw2;
}
// frequency-independent state of one H2O / O2 line at one level-point
// Fast pair form used by both species.  With u = f^2, fc = line centre, w = half width:
//   D1 D2 = ((f-fc)^2 + w^2)((f+fc)^2 + w^2) = (u - fc^2 + w^2)^2 + 4 fc^2 w^2 = vv^2 + c4,  vv = u - c3
//   D1 + D2 = 2 (vv + 2 fc^2),   d1 D2 - d2 D1 = 2 fc (vv - 2 w^2)
// so a line's two wings are ONE fraction (n1*vv + n0)/(vv^2 + c4) per frequency, and two LINES a, b
// share one reciprocal:  na/Da + nb/Db = (na Db + nb Da)/(Da Db):
//   2 DADD (vv) + 2 DFMA (D) + 2 DFMA (n) + DMUL (Da Db) + DMUL, DFMA (numerator) + reciprocal
//   (MUFU.RCP64H + 2 DFMA, rcp_nr1) + DFMA (accumulate) = 12 FP64-pipe instructions and 1 MUFU per
//   (line pair, frequency), i.e. 6 per line instead of 7, with half the MUFU and f^2 loads.
// Rounding: the only cancellation is in vv (relative error ~ eps*f/(2*max(|f-fc|, w))); lines that are
// both narrower than 1e-3*fc and closer than 1e-3*fc to a frequency of the block get the exact
// per-wing form as a correction in the slow-path loop.
struct H2OLine {
  double w, wsq, base, shift, sp, w2, dl2, c3, c4, n1, n0, b32;   // b32 = c3 - fl^2 (fp32 mode)
};
struct O2Line {
  double fcen, df, df2, sp, y, gfac, c3, c4, n1, n0, b32;        // b32 = c3 - f_k^2
};
struct PairHot {   // what the pair loop keeps live: 8 doubles
  double c3a, c4a, n1a, n0a, c3b, c4b, n1b, n0b;
};

// The innermost loop of kernel 1: one line PAIR against the FB frequencies of the block (see the
// derivation above H2OLine).  The compiler software-pipelines it two frequencies deep.  Forcing more
// chains in flight (stage-by-stage groups of 4 / 7 / 14 frequencies with ordered FP64 ops) was measured
// on B200 and did NOT help: 0.808 / 0.811 / 0.842 ms against 0.805 ms for config 2 and 8.84 against
// 8.32 ms at F = 700 (profiles/r02c_variants.txt) -- ptxas re-serialises most of it and the extra live
// registers cost spills; neither did 5 or 3 CTAs per SM (0.875 / 0.816 ms).
template <int FB>
__device__ __forceinline__ void pair_accumulate(double (&acc)[FB], const double2* __restrict__ fsq2,
                                                const PairHot& h) {
#pragma unroll
  for (int j = 0; j < FB; ++j) {
    const double u = (j & 1) ? fsq2[j >> 1].y : fsq2[j >> 1].x;   // LDS.128: two frequencies per load
    const double va = u - h.c3a, vb = u - h.c3b;
    const double Da = fma(va, va, h.c4a), Db = fma(vb, vb, h.c4b);
    const double na = fma(h.n1a, va, h.n0a), nb = fma(h.n1b, vb, h.n0b);
    const double num = fma(nb, Da, na * Db);
    acc[j] = fma(num, rcp_nr1(Da * Db), acc[j]);
  }
}

// shared-memory line tables are LINE-major (one row per line, read with LDS.128 at immediate
// offsets): the model's columns, then 1/f_line^2, padded to an even stride
static_assert(kHS % 2 == 0 && kOS % 2 == 0, "row strides must keep 16-byte alignment");
constexpr int kSlowNone = 0, kSlowCorr = 1, kSlowExact = 2;

// ---- far lines on a coarse grid ------------------------------------------------------------------
// A line whose centre is far from the block's frequencies (in units of the block's own width) is a very
// smooth function of u = f^2 across the block: its poles sit at u = c3 +/- i sqrt(c4), a distance
// |c3 - u0| >= rho * h away from the block [u0 - h, u0 + h].  The SUM of all such lines is evaluated
// with the same pair loop at kNQ Chebyshev nodes of the block instead of at its FB frequencies and
// interpolated once per level-point (Lagrange weights, tabulated per block): kNQ instead of FB
// evaluations per far line.  Interpolation error ~2 (rho + sqrt(rho^2-1))^-kNQ of the far lines' own
// contribution; kFarRho is chosen so that this is ~1e-11 (measured against the full evaluation: <= 3e-11
// of the line sum, tests/test_kernel_algebra.py; the parity gate on absorption is 1e-9).  The decision
// is per frequency block (k_abs_blocks) from the table centres, with margins for the pressure shift (the
// 1 GHz the fast form allows; the O2 mixing shift bound) and the width (w^2 <= 37 GHz^2).  A 28- or
// 32-wide block of a sub-band set or dense spectrum sees nearly every line as far (config 3 K-band: 13 of
// 14 H2O, 49 of 49 O2; 50-60 GHz blocks: 23-32 O2 lines); the 14 HATPRO channels as ONE block only 23 of 65.
#ifndef MWRT_ABS_NQ
#define MWRT_ABS_NQ 6
#endif
constexpr int kNQ = MWRT_ABS_NQ;
static_assert(kNQ == 4 || kNQ == 6 || kNQ == 8, "node count: even (rows are read with LDS.128)");
constexpr double kFarRho = kNQ == 4 ? 282.0 : (kNQ == 6 ? 34.1 : 11.9);   // 10^(11/kNQ) = rho + sqrt(rho^2 - 1)
#ifndef MWRT_ABS_FAR_BUILD
#define MWRT_ABS_FAR_BUILD 1   // 0 compiles the far-line code out (tuning experiments)
#endif
#ifndef MWRT_ABS_FAR_MINFB
#define MWRT_ABS_FAR_MINFB 17  // narrowest block that carries the far-line code
#endif
// The far-line code is compiled into the WIDE-block kernels only (28/32 frequencies per block: dense
// spectra and sub-band sets, where nearly every line is far: config 3 kernel 1 -27 %, config 4 -30 %).
// For the 14 HATPRO channels as ONE block only 23 of 65 lines are far and what the coarse grid saves
// (7 % of the FP64 work) is lost again to the latency-bound hoists of the far pairs and to ptxas'
// schedule of the near loops (measured 1.155 vs 1.140 ms per config-2 step), so the narrow kernels keep
// the lean code.
template <int FB>
__host__ __device__ constexpr bool far_built() { return MWRT_ABS_FAR_BUILD != 0 && FB >= MWRT_ABS_FAR_MINFB; }
// c_q = cos((2q+1) pi / (2 kNQ)) and 1 / prod_{q != n} (c_n - c_q)
#if MWRT_ABS_NQ == 4
static __device__ const double kChebNode[kNQ] = {0.9238795325112867, 0.38268343236508984, -0.3826834323650897,
                                                 -0.9238795325112867};
static __device__ const double kChebInvC[kNQ] = {0.7653668647301797, -1.8477590650225737, 1.8477590650225735,
                                                 -0.7653668647301796};
#elif MWRT_ABS_NQ == 6
static __device__ const double kChebNode[kNQ] = {0.9659258262890683,   0.7071067811865476,  0.25881904510252074,
                                                 -0.25881904510252063, -0.7071067811865475, -0.9659258262890682};
static __device__ const double kChebInvC[kNQ] = {1.3803682405467776, -3.771236166328254, 5.151604406875032,
                                                 -5.151604406875032, 3.771236166328254,  -1.3803682405467776};
#else
static __device__ const double kChebNode[kNQ] = {0.9807852804032304,  0.8314696123025452, 0.5555702330196023,
                                                 0.19509032201612833, -0.1950903220161282, -0.555570233019602,
                                                 -0.8314696123025453, -0.9807852804032304};
static __device__ const double kChebInvC[kNQ] = {3.121445152258054,  -8.889123728313642,  13.30351379684073,
                                                 -15.69256448645169, 15.692564486451694,  -13.303513796840717,
                                                 8.889123728313624,  -3.1214451522580533};
#endif

__host__ __device__ inline int abs_smem_doubles(int nh, int no, int FB) {
  const int nop = (no + 1) & ~1;
  // hrow, orow, fblk, fsq, unode, mint, dfm, dfp, dminh, dmino, then ints: hcls[nh], hord[nh], cnt[4],
  // hchk[nh], ochk[nop], oord[nop]
  const int n = nh * kHS + nop * kOS + 2 * FB + kNQ + FB * kNQ + 2 * nh * FB + nh + nop +
                (3 * nh + 2 * nop + 4 + 1) / 2 + 1;
  return (n + 1) & ~1;   // even: what follows (the fp32 tables) stays 16-byte aligned
}

#ifndef MWRT_ABS_MINB
#define MWRT_ABS_MINB 4
#endif
#ifndef MWRT_ABS_O2_PIPELINE
#define MWRT_ABS_O2_PIPELINE 1
#endif
#ifndef MWRT_ABS_WIDE_DEFAULT
#define MWRT_ABS_WIDE_DEFAULT 1
#endif
// F32 = fp32 line-shape mode (mwrt_plan_set_precision): the two innermost pair loops run on the FP32
// pipe; everything hoisted per (level-point, line), the corrections (speed dependence, narrow lines),
// the continua and the accumulation across lines stay fp64.  To keep the one cancellation of the pair
// form out of fp32, vv = f^2 - c3 is formed as (f^2 - fl^2) - (c3 - fl^2): the first difference is
// tabulated per frequency block in fp64 and rounded once, the second is a small per-line quantity.  Strengths are
// scaled by 2^40 so that no term comes near the fp32 denormal range (undone exactly at the end).
constexpr double kF32Scale = 1099511627776.0, kF32Unscale = 1.0 / 1099511627776.0;   // 2^40
struct PairHotF {
  float ba, c4a, n1a, n0a, bb, c4b, n1b, n0b;
};
// fp32 form of the pair loop: the detunings (f^2 - f_line^2) come four at a time from the CTA's fp32
// tables (LDS.128), everything else lives in registers
template <int FB>
__device__ __forceinline__ void pair_accumulate_f32(float (&accf)[FB], const float* __restrict__ Aa,
                                                    const float* __restrict__ Ab, const PairHotF& h) {
  const float4* Aa4 = reinterpret_cast<const float4*>(Aa);
  const float4* Ab4 = reinterpret_cast<const float4*>(Ab);
#pragma unroll
  for (int j4 = 0; j4 < FB; j4 += 4) {
    const float4 qa = Aa4[j4 >> 2], qb = Ab4[j4 >> 2];
    const float ua[4] = {qa.x, qa.y, qa.z, qa.w}, ub[4] = {qb.x, qb.y, qb.z, qb.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (j4 + q < FB) {
        const float va = ua[q] - h.ba, vb = ub[q] - h.bb;
        const float Da = fmaf(va, va, h.c4a), Db = fmaf(vb, vb, h.c4b);
        const float na = fmaf(h.n1a, va, h.n0a), nb = fmaf(h.n1b, vb, h.n0b);
        const float num = fmaf(nb, Da, na * Db);
        accf[j4 + q] = fmaf(num, rcp_f32(Da * Db), accf[j4 + q]);
      }
    }
  }
}

// doubles of one frequency block's scratch (everything behind the line tables, the fp32 tables included)
__host__ __device__ inline int abs_block_doubles(int nh, int no, int FB, bool f32) {
  const int nop = (no + 1) & ~1;
  const int n = abs_smem_doubles(nh, no, FB) - (nh * kHS + nop * kOS) + (f32 ? (nh + nop) * ((FB + 3) & ~3) / 2 : 0);
  return (n + 1) & ~1;
}

// Per-block scratch, ONE CTA per frequency block, once per (plan, knobs): the block's frequencies and
// their squares, the Chebyshev nodes and Lagrange weights of the far-line grid, the exact detunings for
// the slow paths, the line classes and processing orders, the fp32 detuning tables.  Kernel 1's CTAs
// copy their block's record instead of rebuilding it.  Round 1 rebuilt it in each of 7 813 CTAs per
// launch.  Measured on B200 (profiles/r02t_far_lines_experiments.txt): the per-CTA prologue cost 4 % of
// kernel 1 in that form (config 2: 0.776 -> 0.748 ms with the copy) and 12-14 % once the far-line tables
// were added to it, although it runs once per CTA -- more than its own instructions explain; the
// likely mechanism is its code (~2 000 instructions) passing through the 32 KB instruction cache under
// the resident CTAs' hot loops every time a CTA starts.
template <int FB, bool F32>
__global__ void __launch_bounds__(kAbsThreads)
k_abs_blocks(const ModelDev m, const int F, const double* __restrict__ frq, const int far_on,
             double* __restrict__ blocks, const int blk_stride) {
  extern __shared__ __align__(16) double smem[];
  const int nh = m.n_h2o, no = m.n_o2;
  const int nop = (no + 1) & ~1;
  double* hrow = smem;
  double* orow = hrow + nh * kHS;
  double* fblk = orow + nop * kOS;
  double* fsq = fblk + FB;
  double* unode = fsq + FB;
  double* mint = unode + kNQ;
  double* dfm = mint + FB * kNQ;
  double* dfp = dfm + nh * FB;
  double* dminh = dfp + nh * FB;
  double* dmino = dminh + nh;
  int* hcls = reinterpret_cast<int*>(dmino + nop);
  int* hord = hcls + nh;
  int* cnt = hord + nh;
  int* hchk = cnt + 4;
  int* ochk = hchk + nh;
  int* oord = ochk + nop;
  constexpr int FBP = (FB + 3) & ~3;
  float* a32h = reinterpret_cast<float*>(smem + abs_smem_doubles(nh, no, FB));
  float* a32o = a32h + nh * FBP;
  for (int i = threadIdx.x; i < m.n_rows; i += blockDim.x) hrow[i] = m.rows[i];
  for (int i = threadIdx.x; i < blk_stride; i += blockDim.x) fblk[i] = 0.0;   // padding is defined
  const int fb = blockIdx.x, f0 = fb * FB;
  __syncthreads();
  {
    if (threadIdx.x < FB) {
      int jf = f0 + threadIdx.x;
      const double f = frq[jf < F ? jf : F - 1];  // tail block: repeat the last frequency
      fblk[threadIdx.x] = f;
      fsq[threadIdx.x] = f * f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nh * FB; i += blockDim.x) {
      double fl = hrow[(i / FB) * kHS + MWRT_H2O_FL], f = fblk[i % FB];
      dfm[i] = f - fl;
      dfp[i] = f + fl;
    }
    if (F32) {
      for (int i = threadIdx.x; i < nh * FBP; i += blockDim.x) {
        const int jj = i % FBP;
        const double fl = hrow[(i / FBP) * kHS + MWRT_H2O_FL], f = fblk[jj < FB ? jj : FB - 1];
        a32h[i] = (float)((f - fl) * (f + fl));
      }
      for (int i = threadIdx.x; i < nop * FBP; i += blockDim.x) {
        const int jj = i % FBP;
        const double fk = orow[(i / FBP) * kOS + MWRT_O2_F], f = fblk[jj < FB ? jj : FB - 1];
        a32o[i] = (float)((f - fk) * (f + fk));
      }
    }
    for (int i = threadIdx.x; i < nh; i += blockDim.x) {
      double fl = hrow[i * kHS + MWRT_H2O_FL];
      double dmin = fabs(fblk[0] - fl), lo = fblk[0], hi = fblk[0];
      for (int j = 1; j < FB; ++j) {
        dmin = fmin(dmin, fabs(fblk[j] - fl));
        lo = fmin(lo, fblk[j]);
        hi = fmax(hi, fblk[j]);
      }
      dminh[i] = dmin;
      // Per-CTA line classes w.r.t. the 750 GHz cutoff, decided from the table centre with a 1 GHz
      // margin for the pressure shift:
      //   kAllIn    both wings of every block frequency inside  -> fast pair form
      //   kAllOut   both wings outside (e.g. 916 GHz below 166 GHz) -> contributes exactly 0: skipped
      //   anything else (one wing out, or a cutoff crossing inside the block) -> exact per-wing loop
      const bool w0_in = (fl - lo < 749.0) && (hi - fl < 749.0);   // |f - fc| < 750 for every f
      int cls = kMixed;
      if (w0_in && fl + hi < 749.0) cls = kAllIn;
      else if (w0_in && fl + lo > 751.0) cls = kWing1Out;
      else if (fl - hi > 751.0 || lo - fl > 751.0) cls = kAllOut;
      // Corrections to the fast form are possible only for speed-dependent lines and for lines within
      // 1e-3*fl (+ the 1 GHz shift margin of the fast path) of a block frequency: decided here once
      // per block, so the per-thread tests run for those lines only (HATPRO: the 22 GHz line).
      const bool sd_cap = (m.h2o_sd != 0) && (hrow[i * kHS + MWRT_H2O_W2] != 0.0 || hrow[i * kHS + MWRT_H2O_W2S] != 0.0);
      const int chk = (sd_cap || dmin < 1e-3 * fl + 1.0) ? 1 : 0;
      hchk[i] = chk;
      // far line (see kNQ): all-in, never corrected, and its poles at least kFarRho block half-widths
      // away from the block centre in u = f^2 whatever the shift (< 1 GHz) and width (w^2 < 37)
      const double u_lo = lo * lo, u_hi = hi * hi, u0 = 0.5 * (u_lo + u_hi);
      const double hh = fmax(0.5 * (u_hi - u_lo), 1e-3 * u0);
      if (far_built<FB>() && far_on && !F32 && cls == kAllIn && !chk && fabs(fl * fl - u0) - 2.0 * fl - 37.0 >= kFarRho * hh)
        cls = kFarIn;
      hcls[i] = cls;
    }
    for (int i = threadIdx.x; i < nop; i += blockDim.x) {
      double fk = orow[i * kOS + MWRT_O2_F];
      double dmin = fabs(fblk[0] - fk), lo = fblk[0], hi = fblk[0];
      for (int j = 1; j < FB; ++j) {
        dmin = fmin(dmin, fabs(fblk[j] - fk));
        lo = fmin(lo, fblk[j]);
        hi = fmax(hi, fblk[j]);
      }
      dmino[i] = dmin;
      // a line can be 'narrow and very close' only within 1e-3*fk + its largest mixing shift
      // (|dnu0| + |dnu1| times a bound of 4 on den^2 and |th1| for atmospheric T, p) of a frequency
      const double shmax = 4.0 * (fabs(orow[i * kOS + MWRT_O2_DNU0]) + fabs(orow[i * kOS + MWRT_O2_DNU1])) + 1e-6;
      int chk = ((i == 0 && m.o2_sd118) || dmin < 1e-3 * fk + shmax) ? 1 : 0;
      const double u_lo = lo * lo, u_hi = hi * hi, u0 = 0.5 * (u_lo + u_hi);
      const double hh = fmax(0.5 * (u_hi - u_lo), 1e-3 * u0);
      // far O2 line: marked 2 until the ordering below (the padding row beyond the model's lines stays near)
      if (far_built<FB>() && far_on && !F32 && !chk && i < no && fabs(fk * fk - u0) - 2.0 * fk * shmax - 37.0 >= kFarRho * hh) chk = 2;
      ochk[i] = chk;
    }
    // Chebyshev nodes of the block in u and the Lagrange weights from the nodes to the block's
    // frequencies, in the normalised coordinate t = (u - u0)/h: L_n(t) = prod_{q != n} (t - c_q) / C_n with
    // the node cosines c_q and the constants 1/C_n tabulated (no cos, no division, no extra barrier)
    if (far_built<FB>() && !F32) {
      double lo = fblk[0], hi = fblk[0];
      for (int j = 1; j < FB; ++j) {
        lo = fmin(lo, fblk[j]);
        hi = fmax(hi, fblk[j]);
      }
      const double u_lo = lo * lo, u_hi = hi * hi, u0 = 0.5 * (u_lo + u_hi);
      const double hh = fmax(0.5 * (u_hi - u_lo), 1e-3 * u0);
      const double ihh = rcp_fast(hh);
      if (threadIdx.x < kNQ) unode[threadIdx.x] = fma(hh, kChebNode[threadIdx.x], u0);
      for (int i = threadIdx.x; i < FB * kNQ; i += blockDim.x) {
        const int n = i % kNQ;
        const double tt = (fsq[i / kNQ] - u0) * ihh;
        double w = kChebInvC[n];
#pragma unroll
        for (int q = 0; q < kNQ; ++q) w *= (q != n) ? tt - kChebNode[q] : 1.0;
        mint[i] = w;
      }
    }
    __syncthreads();
    // Processing orders, built with warp ballots (a serial loop by one thread would hold the whole CTA
    // at the barrier below for ~100 dependent shared-memory round trips).
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    if ((threadIdx.x >> 5) == 0) {
      // H2O: far lines (an even number, at most 64: an odd one out is simply treated as near), the
      // other all-in lines, then the lines that need the exact loop; all-out lines are dropped
      int nfar_tot = 0, nall_raw = 0;
      for (int b = 0; b < nh; b += 32) {
        const int cls = (b + lane < nh) ? hcls[b + lane] : kAllOut;
        nfar_tot += __popc(__ballot_sync(0xffffffffu, cls == kFarIn));
        nall_raw += __popc(__ballot_sync(0xffffffffu, cls == kAllIn));
      }
      const int nf = nfar_tot > 64 ? 64 : (nfar_tot & ~1);
      const int n_in = nall_raw + nfar_tot;
      int fb_ = 0, ab_ = 0, ob_ = 0;
      for (int b = 0; b < nh; b += 32) {
        const int i = b + lane;
        const int cls = (i < nh) ? hcls[i] : kAllOut;
        const bool isfar = cls == kFarIn;
        const unsigned mf = __ballot_sync(0xffffffffu, isfar);
        const int fpos = fb_ + __popc(mf & lt);
        const bool kept = isfar && fpos < nf;
        const bool isin = (cls == kAllIn) || (isfar && !kept);
        const unsigned mi = __ballot_sync(0xffffffffu, isin);
        const bool isoth = (cls == kWing1Out) || (cls == kMixed);
        const unsigned mo = __ballot_sync(0xffffffffu, isoth);
        if (kept) {
          hord[fpos] = i;
        } else if (isin) {
          hord[nf + ab_ + __popc(mi & lt)] = i;
          if (isfar) hcls[i] = kAllIn;
        } else if (isoth) {
          hord[n_in + ob_ + __popc(mo & lt)] = i;
        }
        fb_ += __popc(mf);
        ab_ += __popc(mi);
        ob_ += __popc(mo);
      }
      if (lane == 0) {
        cnt[2] = nf;
        cnt[0] = n_in;
        cnt[1] = n_in + ob_;
      }
    } else if ((threadIdx.x >> 5) == 1) {
      // O2: far lines (even count) first, then the rest in table order (the padding row last)
      int nfar_tot = 0;
      for (int b = 0; b < nop; b += 32)
        nfar_tot += __popc(__ballot_sync(0xffffffffu, (b + lane < nop) && ochk[b + lane] == 2));
      const int nf = nfar_tot & ~1;
      int fb_ = 0;
      for (int b = 0; b < nop; b += 32) {
        const int i = b + lane;
        const bool isfar = (i < nop) && ochk[i] == 2;
        const unsigned mf = __ballot_sync(0xffffffffu, isfar);
        const int fpos = fb_ + __popc(mf & lt);   // far lines before line i
        if (i < nop) {
          if (isfar && fpos < nf) {
            oord[fpos] = i;
          } else {
            oord[nf + i - (fpos < nf ? fpos : nf)] = i;   // i lines before, min(fpos, nf) of them kept far
            if (isfar) ochk[i] = 0;
          }
        }
        fb_ += __popc(mf);
      }
      if (lane == 0) cnt[3] = nf;
    }
    __syncthreads();
  }
  double* dst = blocks + (size_t)fb * blk_stride;
  for (int i = threadIdx.x; i < blk_stride; i += blockDim.x) dst[i] = fblk[i];
}

template <int FB, bool F32>
__global__ void __launch_bounds__(kAbsThreads, FB > 16 ? 3 : MWRT_ABS_MINB)
k_absorption(const ModelDev m, const long long npts, const int L, const int F, const int nfb,
             const double* __restrict__ p_in, const double* __restrict__ tk_in,
             const double* __restrict__ e_in, const int e_is_rh, double* __restrict__ e_out,
             double* __restrict__ rho_out, const double* __restrict__ frq,
             double* __restrict__ awet, double* __restrict__ adry, int* __restrict__ status,
             const double* __restrict__ blocks, const int blk_stride) {
  extern __shared__ __align__(16) double smem[];
  const int nh = m.n_h2o, no = m.n_o2;
  const int nop = (no + 1) & ~1;          // O2 lines padded to an even count (dummy: strength 0)
  double* hrow = smem;                    // nh x kHS
  double* orow = hrow + nh * kHS;         // nop x kOS
  double* fblk = orow + nop * kOS;        // FB frequencies of this block
  double* fsq = fblk + FB;                // FB: f_j^2
  double* unode = fsq + FB;               // kNQ: Chebyshev nodes of the block in u = f^2 (far lines)
  double* mint = unode + kNQ;             // FB x kNQ: Lagrange weights node -> frequency
  double* dfm = mint + FB * kNQ;          // nh x FB: f_j - fl_i   (slow paths keep the oracle's rounding)
  double* dfp = dfm + nh * FB;            // nh x FB: f_j + fl_i
  double* dminh = dfp + nh * FB;          // nh: min_j |f_j - fl_i|
  double* dmino = dminh + nh;             // nop: min_j |f_j - f_k|
  int* hcls = reinterpret_cast<int*>(dmino + nop);   // nh: H2O line class w.r.t. the 750 GHz cutoff
  int* hord = hcls + nh;                  // nh: processing order, all-in lines first
  int* cnt = hord + nh;                   // [0] all-in lines, [1] lines that contribute at all,
                                          // [2] far H2O lines (even, first in hord), [3] far O2 lines (even)
  int* hchk = cnt + 4;                    // nh: can this line need a correction in this block at all?
  int* ochk = hchk + nh;                  // nop: same for the O2 lines
  int* oord = ochk + nop;                 // nop: O2 processing order, far lines first
  // fp32 line-shape mode: (f_j^2 - f_line^2) rounded to fp32, nh rows then nop rows of FBP floats
  // (rows padded to a multiple of four so that one LDS.128 brings four frequencies)
  constexpr int FBP = (FB + 3) & ~3;
  float* a32h = reinterpret_cast<float*>(smem + abs_smem_doubles(nh, no, FB));
  float* a32o = a32h + nh * FBP;
  __shared__ double etab[kExpTabN];             // exp_tb / log_tb tables (mwrt_common.cuh)
  __shared__ __align__(16) double ltab[128];
  exp_tab_stage(etab);
  log_tab_stage(ltab);
  // line tables: a flat copy of the model's line-major rows (hrow and orow are contiguous)
  for (int i = threadIdx.x; i < m.n_rows; i += blockDim.x) hrow[i] = m.rows[i];
  // Work items = (frequency block, tile of kAbsThreads points); CTA b owns the contiguous item range
  // [items*b/G, items*(b+1)/G) -- normally exactly one item (see launch_abs).
  const long long ntiles = (npts + kAbsThreads - 1) / kAbsThreads;
  const long long nitems_all = ntiles * nfb;
  const long long it_begin = nitems_all * blockIdx.x / gridDim.x;
  const long long it_end = nitems_all * (blockIdx.x + 1) / gridDim.x;
  int cur_fb = -1;
#pragma unroll 1
  for (long long item = it_begin; item < it_end; ++item) {
  const int fb = (int)(item / ntiles);
  const long long tile = item - (long long)fb * ntiles;
  const int f0 = fb * FB;
  // the level-point's inputs are requested before the block scratch is built: their DRAM latency
  // hides under the staging below
  const long long ipt = tile * kAbsThreads + threadIdx.x;
  const bool in_range = ipt < npts;
  const double p_mb = in_range ? p_in[ipt] : 1000.0;
  const double tk_raw = in_range ? tk_in[ipt] : 300.0;
  double e_mb = in_range ? e_in[ipt] : 0.0;
  if (fb != cur_fb) {
    cur_fb = fb;
    __syncthreads();   // everyone is done with the previous block's scratch
    // the block's scratch, built once per launch configuration by k_abs_blocks: a flat 16-byte copy
    const double2* src = reinterpret_cast<const double2*>(blocks + (size_t)fb * blk_stride);
    double2* dst = reinterpret_cast<double2*>(fblk);
    for (int i = threadIdx.x; i < (blk_stride >> 1); i += blockDim.x) dst[i] = src[i];
    __syncthreads();
  }
  if (in_range) {
  const long long pidx = ipt / L;
  const int lev = (int)(ipt - pidx * L);

  // the branch-free exp needs bounded arguments: temperatures outside (50 K, 1000 K) are computed
  // with a stand-in and reported as NaN + status (the reference's realism window is 0..400 K)
  const bool t_ok = (tk_raw > 50.0) && (tk_raw < 1000.0);
  const double tk = t_ok ? tk_raw : 300.0;
  const bool first = (fb == 0);
  if (e_is_rh) e_mb = e_mb * goff_gratch_es(tk, etab, ltab);
  const bool in_nan = isnan(p_mb) || isnan(tk_raw) || isnan(e_mb);
  const bool in_bad = in_nan || !t_ok;
  if (first) {
    if (e_out) e_out[ipt] = in_bad ? nan("") : e_mb;
    if (rho_out) rho_out[ipt] = in_bad ? nan("") : e_mb / ((kRWatVap * 1e-05) * tk);
    if (status && in_bad) atomicOr(&status[pidx], in_nan ? MWRT_ST_NAN_INPUT : MWRT_ST_RANGE);
  }

  double f_lo = fblk[0], f_hi = fblk[0];
#pragma unroll
  for (int j = 1; j < FB; ++j) {
    f_lo = f_lo < fblk[j] ? f_lo : fblk[j];
    f_hi = f_hi > fblk[j] ? f_hi : fblk[j];
  }

  // level-point state, in the rounding order of RTEquation.clearsky_absorption
  const double vx = div_fast(300.0, tk);   // div_fast: <= 1 ulp from IEEE, no slow-path branch
  const double ekpa = div_fast(e_mb, 10.0);
  const double pdrykpa = div_fast(p_mb, 10.0) - ekpa;
  const double t = div_fast(300.0, vx);
  const double pp = (pdrykpa + ekpa) * 10.0;
  const double rho = div_fast(ekpa * 10.0, kRvapAbs * t);

  const long long obase = (pidx * F + f0) * (long long)L + lev;  // + j*L per frequency
  const double2* fsq2 = reinterpret_cast<const double2*>(fsq);   // fsq is 16-byte aligned (even offsets)
  double acc[FB];
  float accf[F32 ? FB : 1];   // fp32 mode: the pair-loop sums of one species (scaled by 2^40); the fp64
                              // `acc` then only collects the corrections
  double tmp[FB];  // slow-path scratch (local memory, dynamically indexed)

  // ------------------------------------------------------------------ H2O (abh2o)
  {
    const double pvap = div_fast(rho * t, m.h2o_pvap_div);
    const double pda = pp - pvap;
    const double den = m.h2o_den_const * rho;
    const double tic = div_fast(m.h2o_reftcon, t);
    const double ticln = log_tb(tic, ltab);
    const double conbase =
        (m.h2o_cf * pda * exp_tb_(m.h2o_xcf * ticln) + m.h2o_cs * pvap * exp_tb_(m.h2o_xcs * ticln)) * pvap;
    const double ti = div_fast(m.h2o_reftline, t);
    const double tiln = log_tb(ti, ltab);
    const double omti = 1.0 - ti;
    const bool sd_on = m.h2o_sd != 0;

    // full line state from its table row R and its three exps (width, self width, strength)
    auto line_state = [&](const double* R, double e_x, double e_xs, double e_s) {
      H2OLine h;
      h.w = R[MWRT_H2O_W0] * pda * e_x + R[MWRT_H2O_W0S] * pvap * e_xs;
      // ti^0 = 1 exactly: most lines have no shift exponent, skip their exp (warp-uniform branch)
      double shf = R[MWRT_H2O_SH] * pda * (1.0 - R[MWRT_H2O_AAIR] * tiln);
      double shs = R[MWRT_H2O_SHS] * pvap * (1.0 - R[MWRT_H2O_ASELF] * tiln);
      const double xh = R[MWRT_H2O_XH], xhs = R[MWRT_H2O_XHS];
      if (xh != 0.0) shf *= exp_tb_(xh * tiln);
      if (xhs != 0.0) shs *= exp_tb_(xhs * tiln);
      h.shift = R[MWRT_H2O_SR] * h.w + shf + shs;
      h.w2 = sd_on ? R[MWRT_H2O_W2] * pda + R[MWRT_H2O_W2S] * pvap : 0.0;
      h.dl2 = sd_on ? R[MWRT_H2O_D2] * pda + R[MWRT_H2O_D2S] * pvap : 0.0;
      h.wsq = h.w * h.w;
      h.base = h.w * rcp_fast(562500.0 + h.wsq);
      // strength with (1/fl)^2 folded in; the common f^2 is applied once at the end
      h.sp = R[MWRT_H2O_S1] * e_s * R[kH_IFL2];
      // fast form: sp*w*(D0+D1)/(D0 D1) = (n1*vv + n0) / (vv^2 + c4)
      const double fc = R[MWRT_H2O_FL] + h.shift;
      const double fc2 = fc * fc;
      h.c3 = fc2 - h.wsq;
      h.c4 = 4.0 * fc2 * h.wsq;
      h.n1 = 2.0 * (h.sp * h.w);
      h.n0 = h.n1 * (2.0 * fc2);
      h.b32 = fma(h.shift, 2.0 * R[MWRT_H2O_FL] + h.shift, -h.wsq);
      return h;
    };
    // does the fast form of this line need a correction at some frequency of the block?
    auto needs_corr = [&](const H2OLine& g, int i, double fl) {
      const double fc = fl + g.shift;
      const bool sd_any = (g.w2 > 0.0) && (f_lo - fc < 10.0 * g.w) && (fc - f_hi < 10.0 * g.w);
      const bool narrow = (g.w < 1e-3 * fl) && (dminh[i] - fabs(g.shift) < 1e-3 * fl);
      return sd_any || narrow;
    };

    if constexpr (F32) {
#pragma unroll
      for (int j = 0; j < FB; ++j) accf[j] = 0.0f;
    }
    double basesum = 0.0;
    const int n_in = cnt[0], n_act = cnt[1];
    const int npairs = (n_in + 1) >> 1;
    const int nitems = npairs + (n_act - n_in);
    // far pairs first (see kNQ): the pair loop at the kNQ nodes of the block, then one interpolation to
    // the block's frequencies, which initialises the running sums
    const int nfar = far_built<FB>() ? cnt[2] >> 1 : 0;
    unsigned far_bad = 0;   // far pairs whose pressure shift broke the 1 GHz margin: exact loop below
    if (far_built<FB>() && !F32 && nfar > 0) {
      double accn[kNQ];
#pragma unroll
      for (int n = 0; n < kNQ; ++n) accn[n] = 0.0;
#pragma unroll 1
      for (int it = 0; it < nfar; ++it) {
        const double* Ra = hrow + hord[2 * it] * kHS;
        const double* Rb = hrow + hord[2 * it + 1] * kHS;
        double ea[6], ev[6];
        ea[0] = Ra[MWRT_H2O_X] * tiln;
        ea[1] = Ra[MWRT_H2O_XS] * tiln;
        ea[2] = fma(2.5, tiln, Ra[MWRT_H2O_B2] * omti);
        ea[3] = Rb[MWRT_H2O_X] * tiln;
        ea[4] = Rb[MWRT_H2O_XS] * tiln;
        ea[5] = fma(2.5, tiln, Rb[MWRT_H2O_B2] * omti);
        exp_tb_vec<6>(ea, ev, etab);
        const H2OLine ga = line_state(Ra, ev[0], ev[1], ev[2]);
        const H2OLine gb = line_state(Rb, ev[3], ev[4], ev[5]);
        if ((fabs(ga.shift) < 1.0) && (fabs(gb.shift) < 1.0)) {
          basesum += 2.0 * ga.sp * ga.base + 2.0 * gb.sp * gb.base;
          const PairHot h{ga.c3, ga.c4, ga.n1, ga.n0, gb.c3, gb.c4, gb.n1, gb.n0};
          pair_accumulate<kNQ>(accn, reinterpret_cast<const double2*>(unode), h);
        } else {
          far_bad |= 1u << it;
        }
      }
      const double2* mi2 = reinterpret_cast<const double2*>(mint);
#pragma unroll
      for (int j = 0; j < FB; ++j) {
        double a = 0.0;
#pragma unroll
        for (int n = 0; n < kNQ; n += 2) {
          const double2 w = mi2[(j * kNQ + n) >> 1];
          a = fma(w.x, accn[n], a);
          a = fma(w.y, accn[n + 1], a);
        }
        acc[j] = a;
      }
    } else {
#pragma unroll
      for (int j = 0; j < FB; ++j) acc[j] = 0.0;
    }
    // work items: pairs of all-in lines (fast pair form + corrections), then the remaining lines one
    // at a time through the exact per-wing loop
#pragma unroll 1
    for (int it = far_bad ? 0 : nfar; it < nitems; ++it) {
      int ia, ib, ma = kSlowNone, mb = kSlowNone;
      if (it < nfar) {
        if (!((far_bad >> it) & 1u)) continue;
        ia = hord[2 * it];
        ib = hord[2 * it + 1];
        ma = mb = kSlowExact;
      } else if (it < npairs) {
        const int q = 2 * it;
        const bool has_b = q + 1 < n_in;
        ia = hord[q];
        ib = has_b ? hord[q + 1] : ia;
        const double* Ra = hrow + ia * kHS;
        const double* Rb = hrow + ib * kHS;
        double ea[6], ev[6];
        ea[0] = Ra[MWRT_H2O_X] * tiln;
        ea[1] = Ra[MWRT_H2O_XS] * tiln;
        ea[2] = fma(2.5, tiln, Ra[MWRT_H2O_B2] * omti);
        ea[3] = Rb[MWRT_H2O_X] * tiln;
        ea[4] = Rb[MWRT_H2O_XS] * tiln;
        ea[5] = fma(2.5, tiln, Rb[MWRT_H2O_B2] * omti);
        exp_tb_vec<6>(ea, ev, etab);   // six chains interleaved
        const H2OLine ga = line_state(Ra, ev[0], ev[1], ev[2]);
        const H2OLine gb = line_state(Rb, ev[3], ev[4], ev[5]);
        const bool margin_ok = (fabs(ga.shift) < 1.0) && (fabs(gb.shift) < 1.0);
        if (margin_ok) {
          const double zb = has_b ? 1.0 : 0.0;   // odd count: the last pair's second line is a dummy
          basesum += 2.0 * ga.sp * ga.base + (2.0 * gb.sp * gb.base) * zb;
          if constexpr (F32) {
            const PairHotF h{(float)ga.b32, (float)ga.c4, (float)(ga.n1 * kF32Scale), (float)(ga.n0 * kF32Scale),
                             (float)gb.b32, (float)gb.c4, (float)(gb.n1 * (zb * kF32Scale)),
                             (float)(gb.n0 * (zb * kF32Scale))};
            pair_accumulate_f32<FB>(accf, a32h + ia * FBP, a32h + ib * FBP, h);
          } else {
          const PairHot h{ga.c3, ga.c4, ga.n1, ga.n0, gb.c3, gb.c4, gb.n1 * zb, gb.n0 * zb};
          pair_accumulate<FB>(acc, fsq2, h);
          }
          if (hchk[ia] | hchk[ib]) {   // uniform: almost every pair skips the per-thread tests
            ma = (hchk[ia] && needs_corr(ga, ia, Ra[MWRT_H2O_FL])) ? kSlowCorr : kSlowNone;
            mb = (has_b && hchk[ib] && needs_corr(gb, ib, Rb[MWRT_H2O_FL])) ? kSlowCorr : kSlowNone;
          }
        } else {   // a pressure shift of 1 GHz or more: never in the atmosphere, but stay exact
          ma = kSlowExact;
          mb = has_b ? kSlowExact : kSlowNone;
        }
      } else {
        ia = ib = hord[n_in + (it - npairs)];
        ma = kSlowExact;
        if (hcls[ia] == kWing1Out) {
          // only the f - fc wing is inside the cutoff (e.g. 752 GHz): sp * (w / (d0^2 + w^2) - base)
          const double* R = hrow + ia * kHS;
          double ea[3], ev[3];
          ea[0] = R[MWRT_H2O_X] * tiln;
          ea[1] = R[MWRT_H2O_XS] * tiln;
          ea[2] = fma(2.5, tiln, R[MWRT_H2O_B2] * omti);
          exp_tb_vec<3>(ea, ev, etab);
          const H2OLine g = line_state(R, ev[0], ev[1], ev[2]);
          if (fabs(g.shift) < 1.0 && !(g.w2 > 0.0)) {
            ma = kSlowNone;
            basesum += g.sp * g.base;
            const double sw = g.sp * g.w, sh = g.shift, wsq = g.wsq;
            const double* dm = dfm + ia * FB;
#pragma unroll
            for (int j = 0; j < FB; ++j) {
              const double d0 = dm[j] - sh;
              acc[j] = fma(sw, rcp_fast(fma(d0, d0, wsq)), acc[j]);
            }
          }
        }
      }
      if ((ma | mb) == kSlowNone) continue;
      // slow paths, one line at a time (single instantiation: this code is large):
      //   kSlowExact  the oracle's per-wing evaluation with the 750 GHz cutoff tests
      //   kSlowCorr   what the oracle adds minus what the fast form contributed, for the frequencies
      //               where the speed-dependent shape applies or the line is narrow and very close
#pragma unroll 1
      for (int s = 0; s < 2; ++s) {
        const int mode = s ? mb : ma;
        if (mode == kSlowNone) continue;
        const int i = s ? ib : ia;
        const double* R = hrow + i * kHS;
        double ea[3], ev[3];
        ea[0] = R[MWRT_H2O_X] * tiln;
        ea[1] = R[MWRT_H2O_XS] * tiln;
        ea[2] = fma(2.5, tiln, R[MWRT_H2O_B2] * omti);
        exp_tb_vec<3>(ea, ev, etab);
        const H2OLine g = line_state(R, ev[0], ev[1], ev[2]);
        const double fl = R[MWRT_H2O_FL];
        const bool narrow = (g.w < 1e-3 * fl) && (dminh[i] - fabs(g.shift) < 1e-3 * fl);
        const double* dm = dfm + i * FB;
        const double* dp = dfp + i * FB;
        if (mode == kSlowCorr && !narrow) {
          // Speed-dependent shape only (22/183 GHz near their centres: every K-band channel in the
          // lower troposphere, so this is not rare).  The line is not narrow, so the fast form
          // equals the Lorentz pair to rounding and the correction is sp * (sd - lorentz) of the
          // f - fc wing.  Two frequencies per step, both shapes inlined in one basic block so their
          // long serial chains (complex sqrt, rational w(z)) interleave.
          const double w10 = 10.0 * g.w;
          const SdLine sdl = sd_prepare(g.w, g.w2, g.dl2);
#pragma unroll 1
          for (int j = 0; j < FB; j += 2) {
            const int j1 = (j + 1 < FB) ? j + 1 : j;
            const double da = dm[j] - g.shift, db = dm[j1] - g.shift;
            const bool ua = fabs(da) < w10, ub = (fabs(db) < w10) && (j1 != j);
            if (!(ua || ub)) continue;
            const double sa = sd_shape_fast(sdl, da);
            const double sb = sd_shape_fast(sdl, db);
            const double la = g.w * rcp_fast(fma(da, da, g.wsq));
            const double lb = g.w * rcp_fast(fma(db, db, g.wsq));
            tmp[j] = ua ? g.sp * (sa - la) : 0.0;
            if (j1 != j) tmp[j1] = ub ? g.sp * (sb - lb) : 0.0;
          }
          // frequencies skipped above must not add stale scratch
#pragma unroll
          for (int j = 0; j < FB; ++j) {
            const double d0 = dm[j] - g.shift;
            acc[j] += (fabs(d0) < w10) ? tmp[j] : 0.0;
          }
          continue;
        }
#pragma unroll 1
        for (int j = 0; j < FB; ++j) {
          const double d0 = dm[j] - g.shift;
          const double d1 = dp[j] + g.shift;
          const bool use_sd = g.w2 > 0.0 && fabs(d0) < 10.0 * g.w;
          if (mode == kSlowExact) {
            double ex = 0.0;
            if (use_sd)
              ex += sd_shape_re(g.w, g.w2, g.dl2, d0) - g.base;
            else if (fabs(d0) < 750.0)
              ex += div_fast(g.w, fma(d0, d0, g.wsq)) - g.base;
            if (fabs(d1) < 750.0) ex += div_fast(g.w, fma(d1, d1, g.wsq)) - g.base;
            tmp[j] = g.sp * ex;
            continue;
          }
          if (!use_sd && !narrow) {   // this frequency keeps the fast form: no correction
            tmp[j] = 0.0;
            continue;
          }
          double ex;   // what the oracle adds for this line: sum over wings of (shape - base)
          if (use_sd)
            ex = sd_shape_re(g.w, g.w2, g.dl2, d0) - g.base;
          else
            ex = div_fast(g.w, fma(d0, d0, g.wsq)) - g.base;
          ex += div_fast(g.w, fma(d1, d1, g.wsq)) - g.base;
          // minus the fast form's contribution (base term included; equal to what the pair loop
          // added up to the rounding of the pair merge)
          const double vv = fsq[j] - g.c3;
          const double fast = fma(g.n1, vv, g.n0) * rcp_fast(fma(vv, vv, g.c4));
          tmp[j] = fma(g.sp, ex + 2.0 * g.base, -fast);
        }
#pragma unroll
        for (int j = 0; j < FB; ++j) acc[j] += tmp[j];
      }
    }
    const double lc = m.h2o_line_const * den;
#pragma unroll
    for (int j = 0; j < FB; ++j) {
      const double f = fblk[j];
      const double lsum = F32 ? fma((double)accf[j], kF32Unscale, acc[j]) : acc[j];
      double a = (f * f) * fma(lc, lsum - basesum, conbase);
      if (rho <= 0.0) a = 0.0;
      if (in_bad) a = nan("");
      if (f0 + j < F) awet[obase + (long long)j * L] = a;
    }
  }

  // ------------------------------------------------------------------ O2 (o2abs) + N2
  {
    const double th = div_fast(300.0, t);
    const double th1 = th - 1.0;
    const double thln = log_tb(th, ltab);
    const double b = exp_tb_(m.o2_x * thln);
    const double preswv = div_fast(rho * t, m.o2_pvap_div);
    const double presda = pp - preswv;
    const double den = 0.001 * (presda * b + m.o2_h2o_broad * preswv * th);
    const double dens = 0.001 * (presda + m.o2_h2o_broad * preswv) * th;
    const double dfnr = m.o2_wb300 * den;
    const double pe2 = (m.o2_mixing == 1) ? 0.0 : den * den;   // first order: g = 1, no line shift
    const double ymul = (m.o2_mixing == 1) ? 0.001 * pp * b : den;
    const bool dens118 = m.o2_dens118 != 0;

    auto hoist = [&](int k, double expv) {   // expv = exp(-be_k * th1)
      const double* R = orow + k * kOS;
      O2Line o;
      const double wden = (dens118 && k == 0) ? dens : den;
      o.df = R[MWRT_O2_W300] * wden;
      o.y = ymul * (R[MWRT_O2_Y0] + R[MWRT_O2_Y1] * th1);
      o.gfac = 1.0 + pe2 * (R[MWRT_O2_G0] + R[MWRT_O2_G1] * th1);
      o.fcen = R[MWRT_O2_F] + pe2 * (R[MWRT_O2_DNU0] + R[MWRT_O2_DNU1] * th1);
      o.sp = R[MWRT_O2_S300] * expv * R[kO_IFK2];
      o.df2 = o.df * o.df;
      // fast form: sp*(sf1+sf2) = (n1*vv + n0)/(vv^2 + c4) with As = sp*df*g, P = sp*y*fc:
      //   n1 = 2 (As + P),  n0 = 4 (As fc^2 - P w^2)
      const double as = o.sp * (o.df * o.gfac);
      const double pp_ = (o.sp * o.y) * o.fcen;
      const double fc2 = o.fcen * o.fcen;
      o.c3 = fc2 - o.df2;
      o.c4 = 4.0 * fc2 * o.df2;
      o.n1 = 2.0 * (as + pp_);
      o.n0 = 4.0 * fma(as, fc2, -pp_ * o.df2);
      const double dnu = o.fcen - R[MWRT_O2_F];
      o.b32 = fma(dnu, 2.0 * R[MWRT_O2_F] + dnu, -o.df2);
      return o;
    };

    if constexpr (F32) {
#pragma unroll
      for (int j = 0; j < FB; ++j) accf[j] = 0.0f;
    }
    // Lines are taken in the block's order `oord` (far lines first, then table order), two at a time.
    // hoisted state of the pair at positions (2q, 2q+1) of that order, software-pipelined loops below
    struct O2Pair { double c3a, c4a, n1a, n0a, ba, c3b, c4b, n1b, n0b, bb; };
    auto pair_hot = [&](int q) {
      const int ka = far_built<FB>() ? oord[2 * q] : 2 * q, kb = far_built<FB>() ? oord[2 * q + 1] : 2 * q + 1;
      double ea[2], ev[2];
      ea[0] = -orow[ka * kOS + MWRT_O2_BE] * th1;
      ea[1] = -orow[kb * kOS + MWRT_O2_BE] * th1;
      exp_tb_vec<2>(ea, ev, etab);
      const O2Line a = hoist(ka, ev[0]), c = hoist(kb, ev[1]);
      return O2Pair{a.c3, a.c4, a.n1, a.n0, a.b32, c.c3, c.c4, c.n1, c.n0, c.b32};
    };
    // far pairs (see kNQ): pair loop at the block's kNQ nodes, pipelined like the narrow-block loop
    // below (the hoist of the next pair in the same basic block), then ONE interpolation to the block's
    // frequencies, which initialises the running sums
    const int nqf = far_built<FB>() ? cnt[3] >> 1 : 0, nqp = nop >> 1;
    if (far_built<FB>() && !F32 && nqf > 0) {
      double accn[kNQ];
#pragma unroll
      for (int n = 0; n < kNQ; ++n) accn[n] = 0.0;
      O2Pair cur = pair_hot(0);
#pragma unroll 1
      for (int q = 0; q < nqf; ++q) {
        const O2Pair nxt = pair_hot(q + 1 < nqf ? q + 1 : q);   // (last pass: recomputed and dropped)
        const PairHot h{cur.c3a, cur.c4a, cur.n1a, cur.n0a, cur.c3b, cur.c4b, cur.n1b, cur.n0b};
        pair_accumulate<kNQ>(accn, reinterpret_cast<const double2*>(unode), h);
        cur = nxt;
      }
      const double2* mi2 = reinterpret_cast<const double2*>(mint);
#pragma unroll
      for (int j = 0; j < FB; ++j) {
        double a = 0.0;
#pragma unroll
        for (int n = 0; n < kNQ; n += 2) {
          const double2 w = mi2[(j * kNQ + n) >> 1];
          a = fma(w.x, accn[n], a);
          a = fma(w.y, accn[n + 1], a);
        }
        acc[j] = a;
      }
    } else {
#pragma unroll
      for (int j = 0; j < FB; ++j) acc[j] = 0.0;
    }
    // rare: speed-dependent 118.75 GHz line, or a very narrow line very close to a frequency -- what the
    // oracle adds minus what the fast form contributed (lines ka, kb of a near pair; full state recomputed)
    auto o2_corrections = [&](int ka, int kb) {
      const double expa = exp_tb_(-orow[ka * kOS + MWRT_O2_BE] * th1);
      const double expb = exp_tb_(-orow[kb * kOS + MWRT_O2_BE] * th1);
      const O2Line oa = hoist(ka, expa);
      const O2Line ob = hoist(kb, expb);
      const double fka = orow[ka * kOS + MWRT_O2_F], fkb = orow[kb * kOS + MWRT_O2_F];
      const bool sda = m.o2_sd118 && ka == 0 && (f_lo - oa.fcen < 10.0 * oa.df) &&
                       (oa.fcen - f_hi < 10.0 * oa.df);
      const bool sdb = m.o2_sd118 && kb == 0 && (f_lo - ob.fcen < 10.0 * ob.df) &&
                       (ob.fcen - f_hi < 10.0 * ob.df);
      const bool nwa = (oa.df < 1e-3 * fka) && (dmino[ka] - fabs(oa.fcen - fka) < 1e-3 * fka);
      const bool nwb = (ob.df < 1e-3 * fkb) && (dmino[kb] - fabs(ob.fcen - fkb) < 1e-3 * fkb);
      const int ca = (sda || nwa) ? 1 : 0, cb = (sdb || nwb) ? 1 : 0;
      if (!(ca | cb)) return;
#pragma unroll 1
      for (int s = 0; s < 2; ++s) {
        if (!(s ? cb : ca)) continue;
        const int kk = s ? kb : ka;
        const bool sd = s ? sdb : sda;
        const O2Line g = hoist(kk, s ? expb : expa);   // full line state, recomputed
#pragma unroll 1
        for (int j = 0; j < FB; ++j) {
          const double d1 = fblk[j] - g.fcen;
          const double d2 = fblk[j] + g.fcen;
          double sf1;
          if (sd && fabs(d1) < 10.0 * g.df)
            sf1 = o2_sd118_sf1(g.df, d1, g.y);
          else
            sf1 = div_fast(g.df * g.gfac + d1 * g.y, fma(d1, d1, g.df2));
          const double sf2 = div_fast(g.df * g.gfac - d2 * g.y, fma(d2, d2, g.df2));
          const double vv = fsq[j] - g.c3;
          const double fast = fma(g.n1, vv, g.n0) * rcp_fast(fma(vv, vv, g.c4));
          tmp[j] = fma(g.sp, sf1 + sf2, -fast);
        }
#pragma unroll
        for (int j = 0; j < FB; ++j) acc[j] += tmp[j];
      }
    };
    // Narrow blocks (FB <= 16): SOFTWARE PIPELINE over the line pairs -- the hoisted parameters of pair
    // k+1 (two strength exps and ~45 dependent FP64 operations: a latency-bound chain) are computed in
    // the same basic block as the frequency loop of pair k (FB independent chains: throughput-bound), so
    // the scheduler fills the chain's stalls with the loop's work; only the 8-10 doubles of the next
    // pair are carried, the rare correction path recomputes its line state.  Measured on B200
    // (profiles/r02j_variants.txt): 0.804 -> 0.776 ms for config 2, bit-identical; with 28/32-wide blocks
    // the loop already hides what there is to hide and the extra live registers cost 5-6 %, and the same
    // pipeline over the H2O pairs (six exps and two line states in flight) is slower everywhere (0.87 ms)
    // -- both keep the plain form.
    constexpr bool kO2Pipe = (FB <= 16) && (MWRT_ABS_O2_PIPELINE != 0);
    if constexpr (kO2Pipe) {
      if (nqf < nqp) {
      O2Pair cur = pair_hot(nqf);
#pragma unroll 1
      for (int q = nqf; q < nqp; ++q) {
        const O2Pair nxt = pair_hot(q + 1 < nqp ? q + 1 : q);   // (last pass: recomputed and dropped)
        const int ka = far_built<FB>() ? oord[2 * q] : 2 * q, kb = far_built<FB>() ? oord[2 * q + 1] : 2 * q + 1;
        if constexpr (F32) {
          const PairHotF h{(float)cur.ba, (float)cur.c4a, (float)(cur.n1a * kF32Scale), (float)(cur.n0a * kF32Scale),
                           (float)cur.bb, (float)cur.c4b, (float)(cur.n1b * kF32Scale), (float)(cur.n0b * kF32Scale)};
          pair_accumulate_f32<FB>(accf, a32o + ka * FBP, a32o + kb * FBP, h);
        } else {
          const PairHot h{cur.c3a, cur.c4a, cur.n1a, cur.n0a, cur.c3b, cur.c4b, cur.n1b, cur.n0b};
          pair_accumulate<FB>(acc, fsq2, h);
        }
        cur = nxt;
        if (!(ochk[ka] | ochk[kb])) continue;   // uniform, decided per block
        o2_corrections(ka, kb);
      }   // pairs
      }
    } else {
    // Wide blocks: the line strengths need one exp each.  They are evaluated G at a time, interleaved
    // (exp_tb_vec), so the FP64 pipe sees G independent chains; the values are handed to the pair
    // loop by rotating a small register array (no dynamic indexing).
    constexpr int G = 6;
    const int n0 = 2 * nqf;
#pragma unroll 1
    for (int k0 = n0; k0 < nop; k0 += G) {
      double ea[G], ev[G];
#pragma unroll
      for (int g = 0; g < G; ++g)
        ea[g] = -orow[(far_built<FB>() ? oord[k0 + g < nop ? k0 + g : nop - 1] : (k0 + g < nop ? k0 + g : nop - 1)) * kOS + MWRT_O2_BE] * th1;
      exp_tb_vec<G>(ea, ev, etab);
#pragma unroll 1
      for (int gi = 0; gi < G; gi += 2) {
        const int k = k0 + gi;
        if (k >= nop) break;
        const int ka = far_built<FB>() ? oord[k] : k, kb = far_built<FB>() ? oord[k + 1] : k + 1;
        const double expa = ev[0], expb = ev[1];
#pragma unroll
        for (int q = 0; q + 2 < G; ++q) ev[q] = ev[q + 2];
        const O2Line oa = hoist(ka, expa);
        const O2Line ob = hoist(kb, expb);
        if constexpr (F32) {
          const PairHotF h{(float)oa.b32, (float)oa.c4, (float)(oa.n1 * kF32Scale), (float)(oa.n0 * kF32Scale),
                           (float)ob.b32, (float)ob.c4, (float)(ob.n1 * kF32Scale), (float)(ob.n0 * kF32Scale)};
          pair_accumulate_f32<FB>(accf, a32o + ka * FBP, a32o + kb * FBP, h);
        } else {
        const PairHot h{oa.c3, oa.c4, oa.n1, oa.n0, ob.c3, ob.c4, ob.n1, ob.n0};
        pair_accumulate<FB>(acc, fsq2, h);
        }
        if (!(ochk[ka] | ochk[kb])) continue;   // uniform, decided per block
        o2_corrections(ka, kb);
      }   // pairs of the group
    }     // groups
    }
    const double scale = m.o2_abs_const * presda * (th * th * th);
    const double pdry_mb = pdrykpa * 10.0;
    const double n2base = m.n2_n * (m.n2_l * pdry_mb * pdry_mb * exp_tb_(m.n2_m * thln));
#pragma unroll
    for (int j = 0; j < FB; ++j) {
      const double f = fblk[j];
      const double f2 = f * f;
      const double nonres = m.o2_nonres * dfnr * rcp_fast(th * fma(dfnr, dfnr, f2));
      const double lsum = F32 ? fma((double)accf[j], kF32Unscale, acc[j]) : acc[j];
      double o2 = scale * (f2 * (nonres + lsum));
      o2 = (o2 < 0.0) ? 0.0 : o2;  // max(o2abs, 0)
      o2 *= m.o2_iso;
      double fdep = 1.0;
      if (m.n2_fdep) {
        const double fr = f * (1.0 / 450.0);
        fdep = 0.5 + 0.5 * rcp_fast(fma(fr, fr, 1.0));
      }
      double a = o2 + n2base * fdep * f2;
      if (in_bad) a = nan("");
      if (f0 + j < F) adry[obase + (long long)j * L] = a;
    }
  }

  }   // in_range
  }   // work items
}

// Stream-ordered scratch for the plan-less stage entry point: a private pool that keeps its memory
// between calls (the default pool returns it at every synchronisation, 0.2 ms per call measured)
static cudaError_t abs_scratch_alloc(double** out, size_t bytes, cudaStream_t st) {
  static std::mutex mu;
  static cudaMemPool_t pools[64] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64) return cudaMallocAsync(out, bytes, st);
  {
    std::lock_guard<std::mutex> lk(mu);
    if (!pools[dev]) {
      cudaMemPoolProps props = {};
      props.allocType = cudaMemAllocationTypePinned;
      props.location.type = cudaMemLocationTypeDevice;
      props.location.id = dev;
      e = cudaMemPoolCreate(&pools[dev], &props);
      if (e != cudaSuccess) return e;
      unsigned long long keep = ~0ull;
      e = cudaMemPoolSetAttribute(pools[dev], cudaMemPoolAttrReleaseThreshold, &keep);
      if (e != cudaSuccess) return e;
    }
  }
  return cudaMallocFromPoolAsync(out, bytes, pools[dev], st);
}

template <int FB, bool F32>
static int launch_abs(const ModelDev& m, int P, int L, int F, const double* p, const double* tk,
                      const double* e, int e_is_rh, double* e_out, double* rho_out, const double* frq,
                      double* awet, double* adry, int* status, AbsBlocks* cache, cudaStream_t st) {
  // MWRT_ABS_FAR=0 evaluates every line at every frequency (the round-1 form); read per call like
  // MWRT_ABS_WIDE so that tests compare the two paths in one process
  const char* fenv = getenv("MWRT_ABS_FAR");
  const int far_on = fenv ? atoi(fenv) : 1;
  const long long npts = (long long)P * L;
  const int nfb = (F + FB - 1) / FB;
  const long long items = ((npts + kAbsThreads - 1) / kAbsThreads) * nfb;
  // One work item per CTA whenever the grid allows it.  Measured on B200 (config 2): a persistent grid
  // (one resident wave looping over items) is 6-10 % SLOWER -- a static split quantises to whole items
  // per CTA and lets the warps of a CTA drift apart (more instruction-cache misses).  The item loop in
  // the kernel only matters beyond 2^31-1 items.
  const unsigned grid = (unsigned)(items < 0x7fffffffLL ? items : 0x7fffffffLL);
  const int nop = (m.n_o2 + 1) & ~1;
  size_t smem = sizeof(double) * (size_t)abs_smem_doubles(m.n_h2o, m.n_o2, FB) +
                (F32 ? sizeof(float) * (size_t)(m.n_h2o + nop) * ((FB + 3) & ~3) : 0);
  // The shared-memory carve-out is asked for explicitly: 4 (3 for wide blocks) resident CTAs need a
  // little more than 64 KiB since the far-line tables were added.
  {
    static const cudaError_t carve = [] {
      const char* cenv = getenv("MWRT_ABS_CARVEOUT");   // percent of the SM's shared memory; -1 = driver default
      const int pct = cenv ? atoi(cenv) : 50;
      return pct < 0 ? cudaSuccess
                     : cudaFuncSetAttribute(k_absorption<FB, F32>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    }();
    if (carve != cudaSuccess) return fail_arg(MWRT_E_STATE, "k_absorption: cannot set the shared-memory carve-out");
  }
  if (smem > 48 * 1024) {
    static bool once = [] {
      return cudaFuncSetAttribute(k_absorption<FB, F32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  96 * 1024) == cudaSuccess &&
             cudaFuncSetAttribute(k_abs_blocks<FB, F32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  96 * 1024) == cudaSuccess;
    }();
    if (!once) return fail_arg(MWRT_E_STATE, "k_absorption: cannot raise the dynamic shared-memory limit");
  }
  // Per-block scratch: built by k_abs_blocks.  A plan keeps it (`cache`): rebuilt only when the block
  // width, precision, far-line switch or frequency list changed, then the stream is synchronised ONCE so
  // that the plan's other streams may read it.  The stage entry point (no plan) builds it per call into
  // a stream-ordered allocation.
  const int stride = abs_block_doubles(m.n_h2o, m.n_o2, FB, F32);
  const size_t need = (size_t)nfb * stride;
  double* blocks = nullptr;
  if (cache) {
    const bool hit = cache->buf && cache->fb == FB && cache->f32 == (int)F32 && cache->far == far_on &&
                     cache->F == F && cache->frq == frq;
    if (!hit) {
      // a REbuild (knob or precision changed on a live plan): earlier launches on the plan's other
      // streams may still read the old record
      if (cache->buf) MWRT_CUDA_OK(cudaDeviceSynchronize());
      if (cache->cap < need) {
        if (cache->buf) MWRT_CUDA_OK(cudaFree(cache->buf));
        cache->buf = nullptr;
        cache->cap = 0;
        MWRT_CUDA_OK(cudaMalloc(&cache->buf, sizeof(double) * need));
        cache->cap = need;
      }
      k_abs_blocks<FB, F32><<<nfb, kAbsThreads, smem, st>>>(m, F, frq, far_on, cache->buf, stride);
      MWRT_CUDA_OK(cudaGetLastError());
      MWRT_CUDA_OK(cudaStreamSynchronize(st));
      cache->fb = FB; cache->f32 = (int)F32; cache->far = far_on; cache->F = F; cache->frq = frq;
    }
    blocks = cache->buf;
  } else {
    MWRT_CUDA_OK(abs_scratch_alloc(&blocks, sizeof(double) * need, st));
    k_abs_blocks<FB, F32><<<nfb, kAbsThreads, smem, st>>>(m, F, frq, far_on, blocks, stride);
  }
  cudaError_t err = cudaGetLastError();
  if (err == cudaSuccess) {
    k_absorption<FB, F32><<<grid, kAbsThreads, smem, st>>>(m, npts, L, F, nfb, p, tk, e, e_is_rh, e_out,
                                                           rho_out, frq, awet, adry, status, blocks, stride);
    err = cudaGetLastError();
  }
  if (!cache) {   // the per-call record goes back to the pool in stream order, launch error or not
    const cudaError_t ferr = cudaFreeAsync(blocks, st);
    if (err == cudaSuccess) err = ferr;
  }
  MWRT_CUDA_OK(err);
  return 0;
}

// frequency-block choice: exact fit for the 14 HATPRO channels (and 14*k sub-band sets), else 16/8;
// long lists take wide blocks.  f32ls selects the fp32 line-shape instantiation.
int absorption_launch(const ModelDev& m, int P, int L, int F, const double* p, const double* tk,
                      const double* e, int e_is_rh, double* e_out, double* rho_out, const double* frq,
                      double* awet, double* adry, int* status, int f32ls, AbsBlocks* cache, cudaStream_t st) {
#define MWRT_ABS_ARGS m, P, L, F, p, tk, e, e_is_rh, e_out, rho_out, frq, awet, adry, status, cache, st
#define MWRT_ABS_GO(FB_) return f32ls ? launch_abs<FB_, true>(MWRT_ABS_ARGS) : launch_abs<FB_, false>(MWRT_ABS_ARGS)
  // Block width for long frequency lists (MWRT_ABS_WIDE=0 switches the 28/32-wide blocks off): the
  // hoisted line parameters (40 % of a 14-wide block's FP64 work) are shared by every frequency of the
  // block.  Measured on B200 (profiles/r02b_*): 50-70-wide blocks at 2 CTAs per SM (255 registers) are
  // SLOWER than 28/32 at 3 CTAs per SM -- config 3 share 84.8 ms against 74.6 ms, config 4 +9 % -- the
  // lost occupancy costs more than the hoist saves; they were dropped.
  const char* wenv = getenv("MWRT_ABS_WIDE");   // read per call: tests and tuning runs switch it
  const int wide = wenv ? atoi(wenv) : MWRT_ABS_WIDE_DEFAULT;
  if (wide && F >= 112) {
    if (F % 32 == 0) MWRT_ABS_GO(32);
    if (F % 28 == 0) MWRT_ABS_GO(28);
    if (F % 16 != 0 && F % 14 != 0) MWRT_ABS_GO(32);
  }
  if (F % 16 == 0) MWRT_ABS_GO(16);
  if (F % 14 == 0) MWRT_ABS_GO(14);
  if (F <= 8) MWRT_ABS_GO(8);
  MWRT_ABS_GO(16);
#undef MWRT_ABS_GO
#undef MWRT_ABS_ARGS
}

// ------------------------------------------------------------------------------------------------
// Stage entry point only: cloud liquid / ice absorption (P,F,L) elementwise, level fastest.  The fused
// path never materialises these arrays (kernel 3 evaluates mwrt_cloud.cuh inline).
// ------------------------------------------------------------------------------------------------
__global__ void k_cloud_abs(const int liq_model, const long long n, const int L, const int F,
                            const double* __restrict__ tk, const double* __restrict__ denliq,
                            const double* __restrict__ denice, const double* __restrict__ frq,
                            double* __restrict__ aliq, double* __restrict__ aice) {
  __shared__ double etab[kExpTabN];
  __shared__ __align__(16) double ltab[128];
  exp_tab_stage(etab);
  log_tab_stage(ltab);
  __syncthreads();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // (p, f, l)
  if (i >= n) return;
  const int lev = (int)(i % L);
  const long long pf = i / L;
  const int jf = (int)(pf % F);
  const long long ipt = (pf / F) * L + lev;
  const double t = tk[ipt], f = frq[jf];
  const bool t_bad = isnan(t) || !(t > 50.0 && t < 1000.0);   // same window as kernel 1
  if (aliq) {
    const double denl = denliq ? denliq[ipt] : 0.0;
    double a = 0.0;
    if (denl > 0.0 || isnan(denl)) a = liq_abs_unit(liq_model, t, f, etab, ltab) * denl;
    aliq[i] = t_bad ? nan("") : a;
  }
  if (aice) {
    const double deni = denice ? denice[ipt] : 0.0;
    double a = 0.0;
    if (deni > 0.0 || isnan(deni)) a = ice_abs(f, deni);
    aice[i] = t_bad ? nan("") : a;
  }
}

int cloud_abs_launch(const ModelDev& m, int P, int L, int F, const double* tk, const double* denliq,
                     const double* denice, const double* frq, double* aliq, double* aice,
                     cudaStream_t st) {
  const long long n = (long long)P * F * L;
  k_cloud_abs<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(m.liq_model, n, L, F, tk, denliq, denice, frq,
                                                          aliq, aice);
  MWRT_CUDA_OK(cudaGetLastError());
  return 0;
}

// RTEquation.vapor as a standalone stage
__global__ void k_vapor(long long n, const double* __restrict__ tk, const double* __restrict__ rh,
                        double* __restrict__ e, double* __restrict__ rho) {
  __shared__ double etab[kExpTabN];
  __shared__ __align__(16) double ltab[128];
  exp_tab_stage(etab);
  log_tab_stage(ltab);
  __syncthreads();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double t = tk[i];
  const bool t_ok = (t > 50.0) && (t < 1000.0);     // same window as kernel 1; NaN otherwise
  double ev = rh[i] * goff_gratch_es(t_ok ? t : 300.0, etab, ltab);
  if (!t_ok) ev = nan("");
  e[i] = ev;
  if (rho) rho[i] = ev / ((kRWatVap * 1e-05) * t);
}

int vapor_launch(long long n, const double* tk, const double* rh, double* e, double* rho,
                 cudaStream_t st) {
  k_vapor<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, tk, rh, e, rho);
  MWRT_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace mwrt
