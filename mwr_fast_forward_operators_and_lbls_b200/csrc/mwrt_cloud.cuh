// mwrt_cloud.cuh -- cloud liquid / ice absorption (RTEquation.cloudy_absorption ->
// LiqAbsModel.liquid_water_absorption), built from the branch-free
// table exp/log and atan2 of mwrt_common.cuh: no libm call, no IEEE slow path, one code path per
// permittivity model.  Shared by kernel 3 (which evaluates it INLINE for the cloudy levels of a
// profile: lanes of a warp are frequencies of one profile, so "this level is in cloud" is
// warp-uniform and no (P,F,L) liquid/ice arrays exist on the fused path) and by the stage entry point
// mwrt_absorption_f64 (k_cloud_abs, elementwise).
#pragma once
#include "mwrt_common.cuh"

namespace mwrt {

// log(a / b) for complex a, b (principal branch): the modulus through ONE real log of |a|^2/|b|^2, the
// argument as atan2 of a*conj(b) (a positive scale factor does not move the argument, so the quotient
// itself is never formed)
__device__ __forceinline__ cplx clog_ratio(cplx a, cplx b, const double* __restrict__ ltab) {
  const double nr = fma(a.re, b.re, a.im * b.im), ni = fma(a.im, b.re, -a.re * b.im);
  const double ma = fma(a.re, a.re, a.im * a.im), mb = fma(b.re, b.re, b.im * b.im);
  return cplx{0.5 * log_tb(ma * rcp_fast(mb), ltab), atan2_nb(ni, nr)};
}

// Rayleigh absorption of cloud liquid per unit density [Np km^-1 / (g m^-3)]:
//   -0.06286 * Im((eps-1)/(eps+2)) * f,  Im((eps-1)/(eps+2)) = 3 Im(eps) / |eps+2|^2
// liq_model 0: Liebe-91 double Debye; 1: Rosenkranz-2015 (dilec12).  tk is clamped to a physical range
// (profiles outside it are flagged by kernel 1 and come out NaN anyway).
static __device__ __noinline__ double liq_abs_unit(int liq_model, double tk_raw, double f,
                                            const double* __restrict__ etab,
                                            const double* __restrict__ ltab) {
  const double tk = (tk_raw > 50.0 && tk_raw < 1000.0) ? tk_raw : 300.0;
  double er, ei;
  if (liq_model == 0) {
    const double theta1 = 1.0 - div_fast(300.0, tk);
    const double eps0 = 77.66 - 103.3 * theta1;
    const double eps1 = 0.0671 * eps0;
    const double eps2 = 3.52;
    const double fp = (316.0 * theta1 + 146.4) * theta1 + 20.2;
    const double fs = 39.8 * fp;
    const double x1 = div_fast(f, fp), x2 = div_fast(f, fs);
    const double r1 = rcp_fast(fma(x1, x1, 1.0)), r2 = rcp_fast(fma(x2, x2, 1.0));   // 1/(1+ix) = (1-ix)/(1+x^2)
    const double a = (eps0 - eps1) * r1, b = (eps1 - eps2) * r2;
    er = a + b + eps2;
    ei = -(a * x1 + b * x2);
  } else {
    const double tc = tk - 273.15;
    const double lnth = log_tb(div_fast(300.0, tk), ltab);
    double xa[7], xv[7];
    xa[0] = 0.05 * lnth;
    xa[1] = 1.47 * lnth;
    xa[2] = 2.11 * lnth;
    xa[3] = 2.31 * lnth;
    xa[4] = -div_fast(tc, 226.45);
    xa[5] = fmax(fmin(-div_fast(651.4728, tc + 133.07), 700.0), -700.0);
    xa[6] = -div_fast(tc, 103.05);
    exp_tb_vec<7>(xa, xv, etab);
    const double kappa0 = -43.7527 * xv[0] + 299.504 * xv[1] - 399.364 * xv[2] + 221.327 * xv[3];
    const double delta = 80.69715 * xv[4];
    const double sd = 1164.023 * xv[5];
    const double deltab = 4.008724 * xv[6];
    const double hd = 0.5 * deltab;
    // kappa0 - delta z/(sd + z), z = i f:  z/(sd+z) = (f^2 + i f sd)/(sd^2 + f^2)
    const double q = rcp_fast(fma(sd, sd, f * f));
    const double kr = kappa0 - delta * (f * f) * q;
    const double ki = -delta * (f * sd) * q;
    const double f1 = 10.46012 + 0.1454962 * tc + 0.063267156 * (tc * tc) + 0.00093786645 * (tc * tc * tc);
    const cplx z1 = cmake(-0.75 * f1, f1), z2 = cmake(-4500.0, 2000.0);
    const cplx cn = clog_ratio(z2, z1, ltab);                       // cnorm (temperature only)
    const double icn = rcp_fast(fma(cn.re, cn.re, cn.im * cn.im));
    // chip = hd log((z-z2)/(z-z1)) / cnorm,  chij = hd log((z-z2*)/(z-z1*)) / cnorm*
    const cplx lp = clog_ratio(cmake(4500.0, f - 2000.0), cmake(0.75 * f1, f - f1), ltab);
    const cplx lj = clog_ratio(cmake(4500.0, f + 2000.0), cmake(0.75 * f1, f + f1), ltab);
    const cplx chip = cscale(cmul(lp, cconj(cn)), hd * icn);
    const cplx chij = cscale(cmul(lj, cn), hd * icn);
    er = kr + (chip.re + chij.re - deltab);
    ei = ki + (chip.im + chij.im);
  }
  const double e2 = er + 2.0;
  const double im_re = 3.0 * ei * rcp_fast(fma(e2, e2, ei * ei));
  return -0.06286 * im_re * f;
}

// ice: ((8.18645 / wavelength[m]) * IWC * 0.000959553) * db2np, in the oracle's rounding order; the
// frequency-only quotient is formed once per thread
__device__ __forceinline__ double ice_coef(double f) {
  const double wave = kLightC / (f * 1e9);
  return 8.18645 / wave;
}
__device__ __forceinline__ double ice_abs_c(double coef, double deni) {
  return (coef * deni * 0.000959553) * kDb2Np;
}
__device__ __forceinline__ double ice_abs(double f, double deni) { return ice_abs_c(ice_coef(f), deni); }

}  // namespace mwrt
