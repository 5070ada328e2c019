// mwrt_absorption.cu -- kernel 1: gas (and cloud) absorption for every profile x frequency x level.
//
// Replaces, for a whole batch at once, what PyRTlib does inside TbCloudRTE.execute() per
// (angle, frequency, level): RTEquation.clearsky_absorption -> H2OAbsModel.h2o_absorption,
// O2AbsModel.o2_absorption, N2AbsModel.n2_absorption, and RTEquation.cloudy_absorption ->
// LiqAbsModel.liquid_water_absorption (SURVEY.md 8(a) rows 8,10,13; Appendix A.2-A.7).  The reference
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
//             warp-uniform broadcasts; (f -/+ f_line) is tabulated per CTA so the detuning keeps the
//             oracle's rounding order.
// Absorption does not depend on the viewing angle: it is computed once and reused for all angles
// (PyRTlib recomputes it per angle).
#include "mwrt_common.cuh"

namespace mwrt {

#define exp_tb_(x) exp_tb((x), etab)   // etab: the kernel's shared-memory exp table

constexpr int kAbsThreads = 128;
constexpr int kAllIn = 0, kAllOut = 1, kWing1Out = 2, kMixed = 3;

// rare-path helpers are kept out of line: the hot loops must stay small enough for the
// instruction caches (the first version inlined them 14x and stalled on instruction fetch)
// Speed-dependent shapes with division-free complex arithmetic (reciprocals via MUFU + Newton, no
// IEEE slow-path calls, no branches): the 22 GHz line takes this path for every K-band channel in the
// lower troposphere, so it is not that rare (profiles/r01d: ~20 % of the kernel before this rewrite).
__device__ __forceinline__ cplx cdiv_f(cplx a, cplx b) {
  const double r = rcp_fast(fma(b.re, b.re, b.im * b.im));
  return cplx{fma(a.re, b.re, a.im * b.im) * r, fma(a.im, b.re, -a.re * b.im) * r};
}
__device__ __forceinline__ double sqrt_f(double x) {  // x > 0, normal
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double h = 0.5 * y, s = x * y;
  double e = fma(-s, h, 0.5);
  s = fma(s, e, s);
  h = fma(h, e, h);
  e = fma(-s, h, 0.5);
  s = fma(s, e, s);
  h = fma(h, e, h);
  return fma(fma(-s, s, x), h, s);
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
// Rayleigh liquid-water absorption per unit density [Np km^-1 / (g m^-3)]
__device__ __noinline__ double liq_abs_unit(int liq_model, double f, double tk) {
  cplx eps;
  if (liq_model == 0) {  // Liebe-91 double Debye
    const double theta1 = 1.0 - 300.0 / tk;
    const double eps0 = 77.66 - 103.3 * theta1;
    const double eps1 = 0.0671 * eps0;
    const double eps2 = 3.52;
    const double fp = (316.0 * theta1 + 146.4) * theta1 + 20.2;
    const double fs = 39.8 * fp;
    eps = cadd(cdiv(cmake(eps0 - eps1, 0.0), cmake(1.0, f / fp)),
               cdiv(cmake(eps1 - eps2, 0.0), cmake(1.0, f / fs)));
    eps.re += eps2;
  } else {  // Rosenkranz (2015) `dilec12`
    const double tc = tk - 273.15;
    const double theta = 300.0 / tk;
    const double lnth = log(theta);
    const double kap0 = -43.7527 * exp(0.05 * lnth) + 299.504 * exp(1.47 * lnth) -
                        399.364 * exp(2.11 * lnth) + 221.327 * exp(2.31 * lnth);
    const double delta = 80.69715 * exp(-tc / 226.45);
    const double sd = 1164.023 * exp(-651.4728 / (tc + 133.07));
    const double deltb = 4.008724 * exp(-tc / 103.05);
    const double hdelta = deltb / 2.0;
    const double f1 = 10.46012 + 0.1454962 * tc + 0.063267156 * (tc * tc) +
                      0.00093786645 * (tc * tc * tc);
    const cplx z1 = cmake(-0.75 * f1, f1);
    const cplx z2 = cmake(-4500.0, 2000.0);
    const cplx cnorm = clog_(cdiv(z2, z1));
    const cplx z = cmake(0.0, f);
    cplx kappa = cmake(kap0, 0.0);
    kappa = csub(kappa, cdiv(cscale(z, delta), cmake(sd, f)));
    cplx chip = cdiv(cscale(clog_(cdiv(csub(z, z2), csub(z, z1))), hdelta), cnorm);
    cplx chij = cdiv(cscale(clog_(cdiv(csub(z, cconj(z2)), csub(z, cconj(z1)))), hdelta),
                     cconj(cnorm));
    eps = cadd(kappa, cmake(chip.re + chij.re - deltb, chip.im + chij.im));
  }
  cplx re = cdiv(cmake(eps.re - 1.0, eps.im), cmake(eps.re + 2.0, eps.im));
  return -0.06286 * re.im * f;
}

// frequency-independent state of one H2O / O2 line at one level-point
// Fast pair form used by both species.  With u = f^2, fc = line centre, w = half width:
//   D1 D2 = ((f-fc)^2 + w^2)((f+fc)^2 + w^2) = (u - fc^2 + w^2)^2 + 4 fc^2 w^2 = vv^2 + c4,  vv = u - c3
//   D1 + D2 = 2 (vv + 2 fc^2),   d1 D2 - d2 D1 = 2 fc (vv - 2 w^2)
// so a line's two wings cost per frequency: 1 DADD (vv), 1 DFMA (D), 1 DFMA (numerator = n1*vv + n0),
// one reciprocal (MUFU.RCP64H + 3 DFMA) and 1 DFMA to accumulate: 7 FP64-pipe instructions.
// Rounding: the only cancellation is in vv (relative error ~ eps*f/(2*max(|f-fc|, w))); lines that are
// both narrower than 1e-3*fc and closer than 1e-3*fc to a frequency of the block take the exact
// per-wing form in the rare-path loop instead.
struct H2OLine {
  double w, wsq, base, shift, sp, w2, dl2, c3, c4, n1, n0;
};
struct O2Line {
  double fcen, df, df2, sp, y, gfac, c3, c4, n1, n0;
};
// what the hot loops keep live per line (two copies: current + prefetched next); everything else is
// recomputed in the rare path so that 14 running sums + two of these fit in 128 registers
struct H2OHot {
  double c3, c4, n1, n0, b2, w, shift, w2;
};
struct O2Hot {
  double c3, c4, n1, n0, df, fcen;
};

#ifndef MWRT_ABS_MINB
#define MWRT_ABS_MINB 4
#endif
#ifndef MWRT_ABS_PIPELINE
#define MWRT_ABS_PIPELINE 0   // 1: prefetch the next H2O line's state (measured slower: registers)
#endif
template <int FB>
__global__ void __launch_bounds__(kAbsThreads, MWRT_ABS_MINB)
k_absorption(const ModelDev m, const long long npts, const int L, const int F,
             const double* __restrict__ p_in, const double* __restrict__ tk_in,
             const double* __restrict__ e_in, const int e_is_rh, double* __restrict__ e_out,
             double* __restrict__ rho_out, const double* __restrict__ denliq,
             const double* __restrict__ denice, const double* __restrict__ frq,
             double* __restrict__ awet, double* __restrict__ adry, double* __restrict__ aliq,
             double* __restrict__ aice, int* __restrict__ status) {
  extern __shared__ double smem[];
  const int nh = m.n_h2o, no = m.n_o2;
  double* tab = smem;                     // n_table doubles: H2O columns then O2 columns
  double* fblk = tab + m.n_table;         // FB frequencies of this block
  double* dfm = fblk + FB;                // nh x FB: f_j - fl_i
  double* dfp = dfm + nh * FB;            // nh x FB: f_j + fl_i
  double* ifl2 = dfp + nh * FB;           // nh: 1/fl_i^2
  double* ifk2 = ifl2 + nh;               // no: 1/f_k^2
  double* fsq = ifk2 + no;                // FB: f_j^2
  double* dminh = fsq + FB;               // nh: min_j |f_j - fl_i|
  double* dmino = dminh + nh;             // no: min_j |f_j - f_k|
  int* hcls = reinterpret_cast<int*>(dmino + no);   // nh: H2O line class w.r.t. the 750 GHz cutoff
  __shared__ double etab[64];               // exp_tb / log_tb tables (mwrt_common.cuh)
  __shared__ __align__(16) double ltab[128];
  exp_tab_stage(etab);
  log_tab_stage(ltab);
  const int f0 = blockIdx.y * FB;
  for (int i = threadIdx.x; i < m.n_table; i += blockDim.x) tab[i] = m.tables[i];
  if (threadIdx.x < FB) {
    int jf = f0 + threadIdx.x;
    fblk[threadIdx.x] = frq[jf < F ? jf : F - 1];  // tail block: repeat the last frequency
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nh * FB; i += blockDim.x) {
    double fl = tab[MWRT_H2O_FL * nh + i / FB], f = fblk[i % FB];
    dfm[i] = f - fl;
    dfp[i] = f + fl;
  }
  for (int i = threadIdx.x; i < nh; i += blockDim.x) {
    double fl = tab[MWRT_H2O_FL * nh + i];
    ifl2[i] = 1.0 / (fl * fl);
    double dmin = fabs(fblk[0] - fl), lo = fblk[0], hi = fblk[0];
    for (int j = 1; j < FB; ++j) {
      dmin = fmin(dmin, fabs(fblk[j] - fl));
      lo = fmin(lo, fblk[j]);
      hi = fmax(hi, fblk[j]);
    }
    dminh[i] = dmin;
    const bool w0_in = (fl - lo < 749.0) && (hi - fl < 749.0);   // |f - fc| < 750 for every f
    int cls = kMixed;
    if (w0_in && fl + hi < 749.0) cls = kAllIn;
    else if (w0_in && fl + lo > 751.0) cls = kWing1Out;
    else if (fl - hi > 751.0 || lo - fl > 751.0) cls = kAllOut;
    hcls[i] = cls;
  }
  const double* O = tab + MWRT_H2O_NCOL * nh;
  for (int i = threadIdx.x; i < no; i += blockDim.x) {
    double fk = O[MWRT_O2_F * no + i];
    ifk2[i] = 1.0 / (fk * fk);
    double dmin = fabs(fblk[0] - fk);
    for (int j = 1; j < FB; ++j) dmin = fmin(dmin, fabs(fblk[j] - fk));
    dmino[i] = dmin;
  }
  if (threadIdx.x < FB) fsq[threadIdx.x] = fblk[threadIdx.x] * fblk[threadIdx.x];
  __syncthreads();

  const long long ipt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (ipt >= npts) return;
  const long long pidx = ipt / L;
  const int lev = (int)(ipt - pidx * L);

  const double p_mb = p_in[ipt];
  const double tk_raw = tk_in[ipt];
  // the branch-free exp needs bounded arguments: temperatures outside (50 K, 1000 K) are computed
  // with a stand-in and reported as NaN + status (the reference's realism window is 0..400 K)
  const bool t_ok = (tk_raw > 50.0) && (tk_raw < 1000.0);
  const double tk = t_ok ? tk_raw : 300.0;
  double e_mb = e_in[ipt];
  const bool first = (blockIdx.y == 0);
  if (e_is_rh) e_mb = e_mb * goff_gratch_es(tk);
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
  double acc[FB];
  double tmp[FB];  // rare-path scratch (local memory, dynamically indexed)

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
    const double* H = tab;
    const bool sd_on = m.h2o_sd != 0;

    auto hoist = [&](int i) {
      H2OLine h;
      h.w = H[MWRT_H2O_W0 * nh + i] * pda * exp_tb_(H[MWRT_H2O_X * nh + i] * tiln) +
            H[MWRT_H2O_W0S * nh + i] * pvap * exp_tb_(H[MWRT_H2O_XS * nh + i] * tiln);
      // ti^0 = 1 exactly: most lines have no shift exponent, skip their exp (warp-uniform branch)
      double shf = H[MWRT_H2O_SH * nh + i] * pda * (1.0 - H[MWRT_H2O_AAIR * nh + i] * tiln);
      double shs = H[MWRT_H2O_SHS * nh + i] * pvap * (1.0 - H[MWRT_H2O_ASELF * nh + i] * tiln);
      const double xh = H[MWRT_H2O_XH * nh + i], xhs = H[MWRT_H2O_XHS * nh + i];
      if (xh != 0.0) shf *= exp_tb_(xh * tiln);
      if (xhs != 0.0) shs *= exp_tb_(xhs * tiln);
      h.shift = H[MWRT_H2O_SR * nh + i] * h.w + shf + shs;
      h.w2 = sd_on ? H[MWRT_H2O_W2 * nh + i] * pda + H[MWRT_H2O_W2S * nh + i] * pvap : 0.0;
      h.dl2 = sd_on ? H[MWRT_H2O_D2 * nh + i] * pda + H[MWRT_H2O_D2S * nh + i] * pvap : 0.0;
      h.wsq = h.w * h.w;
      h.base = h.w * rcp_fast(562500.0 + h.wsq);
      // strength with (1/fl)^2 folded in; the common f^2 is applied once at the end
      h.sp = H[MWRT_H2O_S1 * nh + i] * exp_tb_(fma(2.5, tiln, H[MWRT_H2O_B2 * nh + i] * omti)) * ifl2[i];
      // fast pair form: sp*w*(D0+D1)/(D0 D1) = (n1*vv + n0) / (vv^2 + c4)
      const double fc = H[MWRT_H2O_FL * nh + i] + h.shift;
      const double fc2 = fc * fc;
      h.c3 = fc2 - h.wsq;
      h.c4 = 4.0 * fc2 * h.wsq;
      h.n1 = 2.0 * (h.sp * h.w);
      h.n0 = h.n1 * (2.0 * fc2);
      return h;
    };

#pragma unroll
    for (int j = 0; j < FB; ++j) acc[j] = 0.0;
    double basesum = 0.0;
    auto hot = [&](int i) {
      const H2OLine g = hoist(i);
      return H2OHot{g.c3, g.c4, g.n1, g.n0, 2.0 * g.sp * g.base, g.w, g.shift, g.w2};
    };
    // Per-CTA line classes w.r.t. the 750 GHz cutoff (uniform branches, decided from the table
    // centre with a 1 GHz margin for the pressure shift):
    //   kAllIn    both wings of every block frequency inside  -> fast pair form
    //   kAllOut   both wings outside (e.g. 916 GHz below 166 GHz) -> contributes exactly 0: skipped
    //   kWing1Out f+fc always outside, f-fc always inside (e.g. 752 GHz) -> one Lorentzian per frequency
    //   kMixed    anything else -> exact per-wing loop
    for (int i = 0; i < nh; ++i) {
      const int cls = hcls[i];
      if (cls == kAllOut) continue;
      const H2OHot h = hot(i);
      const double fl = tab[MWRT_H2O_FL * nh + i];
      const double fc = fl + h.shift;
      const bool margin_ok = fabs(h.shift) < 1.0;
      if (cls == kAllIn && margin_ok) {
        // 7 FP64-pipe instructions per (line, frequency), see H2OLine
        basesum += h.b2;
#pragma unroll
        for (int j = 0; j < FB; ++j) {
          const double vv = fsq[j] - h.c3;
          const double D = fma(vv, vv, h.c4);
          const double num = fma(h.n1, vv, h.n0);
          acc[j] = fma(num, rcp_fast(D), acc[j]);
        }
        // corrections: the speed-dependent shape near line centre, or a line that is both very
        // narrow and very close to a frequency (exact per-wing form in the oracle's rounding order)
        const bool sd_any = (h.w2 > 0.0) && (f_lo - fc < 10.0 * h.w) && (fc - f_hi < 10.0 * h.w);
        const bool narrow = (h.w < 1e-3 * fl) && (dminh[i] - fabs(h.shift) < 1e-3 * fl);
        if (sd_any || narrow) {
          const H2OLine g = hoist(i);   // full line state, recomputed
          const double* dm = dfm + i * FB;
          const double* dp = dfp + i * FB;
#pragma unroll 1
          for (int j = 0; j < FB; ++j) {
            const double d0 = dm[j] - g.shift;
            const double d1 = dp[j] + g.shift;
            const bool use_sd = g.w2 > 0.0 && fabs(d0) < 10.0 * g.w;
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
            // minus what the fast form contributed (recomputed bit-identically), base term included
            const double vv = fsq[j] - g.c3;
            const double fast = fma(g.n1, vv, g.n0) * rcp_fast(fma(vv, vv, g.c4));
            tmp[j] = fma(g.sp, ex + 2.0 * g.base, -fast);
          }
#pragma unroll
          for (int j = 0; j < FB; ++j) acc[j] += tmp[j];
        }
      } else if (cls == kWing1Out && margin_ok && !(h.w2 > 0.0)) {
        // only the f - fc wing is inside the cutoff: sp * (w / (d0^2 + w^2) - base)
        basesum += 0.5 * h.b2;
        const double sw = 0.5 * h.n1, wsq = h.w * h.w;
        const double* dm = dfm + i * FB;
#pragma unroll
        for (int j = 0; j < FB; ++j) {
          const double d0 = dm[j] - h.shift;
          acc[j] = fma(sw, rcp_fast(fma(d0, d0, wsq)), acc[j]);
        }
      } else {
        // exact per-wing evaluation with the cutoff tests (rare: a line within ~1 GHz of f +- 750)
        const H2OLine g = hoist(i);
        const double* dm = dfm + i * FB;
        const double* dp = dfp + i * FB;
#pragma unroll 1
        for (int j = 0; j < FB; ++j) {
          const double d0 = dm[j] - g.shift;
          const double d1 = dp[j] + g.shift;
          double ex = 0.0;
          if (g.w2 > 0.0 && fabs(d0) < 10.0 * g.w)
            ex += sd_shape_re(g.w, g.w2, g.dl2, d0) - g.base;
          else if (fabs(d0) < 750.0)
            ex += div_fast(g.w, fma(d0, d0, g.wsq)) - g.base;
          if (fabs(d1) < 750.0) ex += div_fast(g.w, fma(d1, d1, g.wsq)) - g.base;
          tmp[j] = g.sp * ex;
        }
#pragma unroll
        for (int j = 0; j < FB; ++j) acc[j] += tmp[j];
      }
    }
    const double lc = m.h2o_line_const * den;
#pragma unroll
    for (int j = 0; j < FB; ++j) {
      const double f = fblk[j];
      double a = (f * f) * fma(lc, acc[j] - basesum, conbase);
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
      O2Line o;
      const double wden = (dens118 && k == 0) ? dens : den;
      o.df = O[MWRT_O2_W300 * no + k] * wden;
      o.y = ymul * (O[MWRT_O2_Y0 * no + k] + O[MWRT_O2_Y1 * no + k] * th1);
      o.gfac = 1.0 + pe2 * (O[MWRT_O2_G0 * no + k] + O[MWRT_O2_G1 * no + k] * th1);
      o.fcen = O[MWRT_O2_F * no + k] + pe2 * (O[MWRT_O2_DNU0 * no + k] + O[MWRT_O2_DNU1 * no + k] * th1);
      o.sp = O[MWRT_O2_S300 * no + k] * expv * ifk2[k];
      o.df2 = o.df * o.df;
      // fast pair form: sp*(sf1+sf2) = (n1*vv + n0)/(vv^2 + c4) with As = sp*df*g, P = sp*y*fc:
      //   n1 = 2 (As + P),  n0 = 4 (As fc^2 - P w^2)
      const double as = o.sp * (o.df * o.gfac);
      const double pp_ = (o.sp * o.y) * o.fcen;
      const double fc2 = o.fcen * o.fcen;
      o.c3 = fc2 - o.df2;
      o.c4 = 4.0 * fc2 * o.df2;
      o.n1 = 2.0 * (as + pp_);
      o.n0 = 4.0 * fma(as, fc2, -pp_ * o.df2);
      return o;
    };

#pragma unroll
    for (int j = 0; j < FB; ++j) acc[j] = 0.0;
    auto hot = [&](int k, double expv) {
      const O2Line g = hoist(k, expv);
      return O2Hot{g.c3, g.c4, g.n1, g.n0, g.df, g.fcen};
    };
    // The line strengths need one exp each: a 17-deep serial chain.  They are evaluated G at a time,
    // interleaved (exp_nb_vec), so the FP64 pipe sees G independent chains instead of one; the
    // values are handed to the line loop by rotating a small register array (no dynamic indexing).
    constexpr int G = 7;
    for (int k0 = 0; k0 < no; k0 += G) {
      double ea[G], ev[G];
#pragma unroll
      for (int g = 0; g < G; ++g) ea[g] = -O[MWRT_O2_BE * no + (k0 + g < no ? k0 + g : no - 1)] * th1;
      exp_tb_vec<G>(ea, ev, etab);
#pragma unroll 1
      for (int gi = 0; gi < G; ++gi) {
      const int k = k0 + gi;
      if (k >= no) break;
      const double expv = ev[0];
#pragma unroll
      for (int q = 0; q + 1 < G; ++q) ev[q] = ev[q + 1];
      const O2Hot o = hot(k, expv);
      // sp*(sf1 + sf2) in the fast pair form: 7 FP64-pipe instructions per (line, frequency)
#pragma unroll
      for (int j = 0; j < FB; ++j) {
        const double vv = fsq[j] - o.c3;
        const double D = fma(vv, vv, o.c4);
        const double num = fma(o.n1, vv, o.n0);
        acc[j] = fma(num, rcp_fast(D), acc[j]);
      }
      // rare path: speed-dependent 118.75 GHz line, or a very narrow line very close to a frequency
      const double fk = O[MWRT_O2_F * no + k];
      const bool sd = m.o2_sd118 && k == 0 && (f_lo - o.fcen < 10.0 * o.df) && (o.fcen - f_hi < 10.0 * o.df);
      const bool narrow = (o.df < 1e-3 * fk) && (dmino[k] - fabs(o.fcen - fk) < 1e-3 * fk);
      if (sd || narrow) {
        const O2Line g = hoist(k, expv);   // full line state, recomputed: this path is rare
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
      }   // lines of the group
    }     // groups
    const double scale = m.o2_abs_const * presda * (th * th * th);
    const double pdry_mb = pdrykpa * 10.0;
    const double n2base = m.n2_n * (m.n2_l * pdry_mb * pdry_mb * exp_tb_(m.n2_m * thln));
#pragma unroll
    for (int j = 0; j < FB; ++j) {
      const double f = fblk[j];
      const double f2 = f * f;
      const double nonres = m.o2_nonres * dfnr * rcp_fast(th * fma(dfnr, dfnr, f2));
      double o2 = scale * (f2 * (nonres + acc[j]));
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

  // ------------------------------------------------------------------ cloud
  if (aliq) {
    const double denl = denliq ? denliq[ipt] : 0.0;
    const bool wet = denl > 0.0 || isnan(denl);
#pragma unroll 1
    for (int j = 0; j < FB; ++j) {
      double a = 0.0;
      if (wet) a = liq_abs_unit(m.liq_model, fblk[j], tk) * denl;
      if (in_bad) a = nan("");
      if (f0 + j < F) aliq[obase + (long long)j * L] = a;
    }
  }
  if (aice) {
    const double deni = denice ? denice[ipt] : 0.0;
#pragma unroll 1
    for (int j = 0; j < FB; ++j) {
      const double f = fblk[j];
      const double wave = kLightC / (f * 1e9);
      double a = 0.0;
      if (deni > 0.0 || isnan(deni)) a = ((8.18645 / wave) * deni * 0.000959553) * kDb2Np;
      if (in_bad) a = nan("");
      if (f0 + j < F) aice[obase + (long long)j * L] = a;
    }
  }
}

template <int FB>
static int launch_abs(const ModelDev& m, int P, int L, int F, const double* p, const double* tk,
                      const double* e, int e_is_rh, double* e_out, double* rho_out,
                      const double* denliq, const double* denice, const double* frq, double* awet,
                      double* adry, double* aliq, double* aice, int* status, cudaStream_t st) {
  const long long npts = (long long)P * L;
  dim3 grid((unsigned)((npts + kAbsThreads - 1) / kAbsThreads), (unsigned)((F + FB - 1) / FB));
  size_t smem = sizeof(double) * (size_t)(m.n_table + 2 * FB + 2 * m.n_h2o * FB + 2 * (m.n_h2o + m.n_o2) + (m.n_h2o + 1) / 2 + 1);
  k_absorption<FB><<<grid, kAbsThreads, smem, st>>>(m, npts, L, F, p, tk, e, e_is_rh, e_out, rho_out,
                                                    denliq, denice, frq, awet, adry, aliq, aice,
                                                    status);
  MWRT_CUDA_OK(cudaGetLastError());
  return 0;
}

// frequency-block choice: exact fit for the 14 HATPRO channels (and 14*k sub-band sets), else 16/8
int absorption_launch(const ModelDev& m, int P, int L, int F, const double* p, const double* tk,
                      const double* e, int e_is_rh, double* e_out, double* rho_out,
                      const double* denliq, const double* denice, const double* frq, double* awet,
                      double* adry, double* aliq, double* aice, int* status, cudaStream_t st) {
#define MWRT_ABS_ARGS m, P, L, F, p, tk, e, e_is_rh, e_out, rho_out, denliq, denice, frq, awet, adry, aliq, aice, status, st
  if (F % 16 == 0) return launch_abs<16>(MWRT_ABS_ARGS);
  if (F % 14 == 0) return launch_abs<14>(MWRT_ABS_ARGS);
  if (F <= 8) return launch_abs<8>(MWRT_ABS_ARGS);
  return launch_abs<16>(MWRT_ABS_ARGS);
#undef MWRT_ABS_ARGS
}

// RTEquation.vapor as a standalone stage
__global__ void k_vapor(long long n, const double* __restrict__ tk, const double* __restrict__ rh,
                        double* __restrict__ e, double* __restrict__ rho) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double t = tk[i];
  double ev = rh[i] * goff_gratch_es(t);
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
