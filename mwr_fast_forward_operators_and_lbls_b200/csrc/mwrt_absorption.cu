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

constexpr int kAbsThreads = 128;

template <int FB>
__global__ void __launch_bounds__(kAbsThreads)
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
  __syncthreads();

  const long long ipt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (ipt >= npts) return;
  const long long pidx = ipt / L;
  const int lev = (int)(ipt - pidx * L);

  const double p_mb = p_in[ipt];
  const double tk = tk_in[ipt];
  double e_mb = e_in[ipt];
  const bool first = (blockIdx.y == 0);
  if (e_is_rh) e_mb = e_mb * goff_gratch_es(tk);
  if (first) {
    if (e_out) e_out[ipt] = e_mb;
    if (rho_out) rho_out[ipt] = e_mb / ((kRWatVap * 1e-05) * tk);
    if (status && (isnan(p_mb) || isnan(tk) || isnan(e_mb))) atomicOr(&status[pidx], MWRT_ST_NAN_INPUT);
  }

  double fmin = fblk[0], fmax = fblk[0];
#pragma unroll
  for (int j = 1; j < FB; ++j) {
    fmin = fmin < fblk[j] ? fmin : fblk[j];
    fmax = fmax > fblk[j] ? fmax : fblk[j];
  }

  // level-point state, in the rounding order of RTEquation.clearsky_absorption
  const double vx = 300.0 / tk;
  const double ekpa = e_mb / 10.0;
  const double pdrykpa = p_mb / 10.0 - ekpa;
  const double t = 300.0 / vx;
  const double pp = (pdrykpa + ekpa) * 10.0;
  const double rho = ekpa * 10.0 / (kRvapAbs * t);

  const long long obase = (pidx * F + f0) * (long long)L + lev;  // + j*L per frequency
  double acc[FB];

  // ------------------------------------------------------------------ H2O (abh2o)
  {
    const double pvap = (rho * t) / m.h2o_pvap_div;
    const double pda = pp - pvap;
    const double den = m.h2o_den_const * rho;
    const double tic = m.h2o_reftcon / t;
    const double ticln = log(tic);
    const double conbase =
        (m.h2o_cf * pda * exp(m.h2o_xcf * ticln) + m.h2o_cs * pvap * exp(m.h2o_xcs * ticln)) * pvap;
    const double ti = m.h2o_reftline / t;
    const double tiln = log(ti);
    const double* H = tab;
#pragma unroll
    for (int j = 0; j < FB; ++j) acc[j] = 0.0;
    double basesum = 0.0;
    for (int i = 0; i < nh; ++i) {
      const double fl = H[MWRT_H2O_FL * nh + i];
      const double w = H[MWRT_H2O_W0 * nh + i] * pda * exp(H[MWRT_H2O_X * nh + i] * tiln) +
                       H[MWRT_H2O_W0S * nh + i] * pvap * exp(H[MWRT_H2O_XS * nh + i] * tiln);
      double shf = H[MWRT_H2O_SH * nh + i] * pda * (1.0 - H[MWRT_H2O_AAIR * nh + i] * tiln);
      const double xh = H[MWRT_H2O_XH * nh + i];
      if (xh != 0.0) shf *= exp(xh * tiln);
      double shs = H[MWRT_H2O_SHS * nh + i] * pvap * (1.0 - H[MWRT_H2O_ASELF * nh + i] * tiln);
      const double xhs = H[MWRT_H2O_XHS * nh + i];
      if (xhs != 0.0) shs *= exp(xhs * tiln);
      const double shift = H[MWRT_H2O_SR * nh + i] * w + shf + shs;
      double w2 = 0.0, dl2 = 0.0;
      if (m.h2o_sd) {
        w2 = H[MWRT_H2O_W2 * nh + i] * pda + H[MWRT_H2O_W2S * nh + i] * pvap;
        dl2 = H[MWRT_H2O_D2 * nh + i] * pda + H[MWRT_H2O_D2S * nh + i] * pvap;
      }
      const double wsq = w * w;
      const double base = w / (562500.0 + wsq);
      // strength with (1/fl)^2 folded in; the common f^2 is applied once at the end
      const double sp = H[MWRT_H2O_S1 * nh + i] *
                        exp(2.5 * tiln + H[MWRT_H2O_B2 * nh + i] * (1.0 - ti)) / (fl * fl);
      const double fc = fl + shift;
      const double* dm = dfm + i * FB;
      const double* dp = dfp + i * FB;
      const bool all_in = !(w2 > 0.0) && (fc + fmax < 750.0) && (fc - fmin < 750.0) &&
                          (fmax - fc < 750.0);
      if (all_in) {
        // both wings inside the 750 GHz cutoff for every frequency of the block:
        //   w/(d0^2+w^2) + w/(d1^2+w^2) = w (D0+D1) / (D0 D1)   -> one reciprocal per (line, f)
        const double sw = sp * w;
        basesum = fma(2.0 * sp, base, basesum);
#pragma unroll
        for (int j = 0; j < FB; ++j) {
          const double d0 = dm[j] - shift;
          const double d1 = dp[j] + shift;
          const double D0 = fma(d0, d0, wsq);
          const double D1 = fma(d1, d1, wsq);
          const double q = (D0 + D1) * rcp_fast(D0 * D1);
          acc[j] = fma(sw, q, acc[j]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < FB; ++j) {
          const double d0 = dm[j] - shift;
          const double d1 = dp[j] + shift;
          double r = 0.0;
          if (w2 > 0.0 && fabs(d0) < 10.0 * w) {
            r += sd_shape(w, w2, dl2, d0).re - base;
          } else if (fabs(d0) < 750.0) {
            r += div_fast(w, fma(d0, d0, wsq)) - base;
          }
          if (fabs(d1) < 750.0) r += div_fast(w, fma(d1, d1, wsq)) - base;
          acc[j] = fma(sp, r, acc[j]);
        }
      }
    }
    const double lc = m.h2o_line_const * den;
#pragma unroll
    for (int j = 0; j < FB; ++j) {
      const double f = fblk[j];
      double a = (f * f) * fma(lc, acc[j] - basesum, conbase);
      if (rho <= 0.0) a = 0.0;
      if (f0 + j < F) awet[obase + (long long)j * L] = a;
    }
  }

  // ------------------------------------------------------------------ O2 (o2abs) + N2
  {
    const double th = 300.0 / t;
    const double th1 = th - 1.0;
    const double thln = log(th);
    const double b = exp(m.o2_x * thln);
    const double preswv = rho * t / m.o2_pvap_div;
    const double presda = pp - preswv;
    const double den = 0.001 * (presda * b + m.o2_h2o_broad * preswv * th);
    const double dens = 0.001 * (presda + m.o2_h2o_broad * preswv) * th;
    const double dfnr = m.o2_wb300 * den;
    const double pe2 = den * den;
    const double ymul = (m.o2_mixing == 1) ? 0.001 * pp * b : den;
    const double* O = tab + MWRT_H2O_NCOL * nh;
#pragma unroll
    for (int j = 0; j < FB; ++j) acc[j] = 0.0;
    for (int k = 0; k < no; ++k) {
      const double fk = O[MWRT_O2_F * no + k];
      double df = O[MWRT_O2_W300 * no + k] * den;
      if (m.o2_dens118 && k == 0) df = O[MWRT_O2_W300 * no] * dens;
      const double y = ymul * (O[MWRT_O2_Y0 * no + k] + O[MWRT_O2_Y1 * no + k] * th1);
      double gfac = 1.0, fcen = fk;
      if (m.o2_mixing != 1) {
        gfac = 1.0 + pe2 * (O[MWRT_O2_G0 * no + k] + O[MWRT_O2_G1 * no + k] * th1);
        fcen = fk + pe2 * (O[MWRT_O2_DNU0 * no + k] + O[MWRT_O2_DNU1 * no + k] * th1);
      }
      const double sp = O[MWRT_O2_S300 * no + k] * exp(-O[MWRT_O2_BE * no + k] * th1) / (fk * fk);
      const double df2 = df * df;
      if (m.o2_sd118 && k == 0) {
#pragma unroll
        for (int j = 0; j < FB; ++j) {
          const double f = fblk[j];
          const double d1 = f - fcen, d2 = f + fcen;
          double sf1;
          if (fabs(d1) < 10.0 * df) {
            const double width2 = 0.076 * df;
            cplx xc = cmake((df - 1.5 * width2) / width2, d1 / width2);
            cplx xrt = csqrt_(xc);
            cplx pxw = cscale(cmul(xrt, dcerror(-xrt.im, xrt.re)), kSqrtPi);
            cplx one_m = cmake(2.0 * (1.0 - pxw.re), -2.0 * pxw.im);
            cplx asd = cmul(cmake(1.0, y), one_m);
            sf1 = asd.re / width2;
          } else {
            sf1 = div_fast(df * gfac + d1 * y, fma(d1, d1, df2));
          }
          const double sf2 = div_fast(df * gfac - d2 * y, fma(d2, d2, df2));
          acc[j] = fma(sp, sf1 + sf2, acc[j]);
        }
      } else {
        // sf1 + sf2 = (n1 D2 + n2 D1) / (D1 D2), strength folded into the numerators
        const double as = sp * (df * gfac);
        const double ys = sp * y;
#pragma unroll
        for (int j = 0; j < FB; ++j) {
          const double f = fblk[j];
          const double d1 = f - fcen;
          const double d2 = f + fcen;
          const double D1 = fma(d1, d1, df2);
          const double D2 = fma(d2, d2, df2);
          const double n1 = fma(d1, ys, as);
          const double n2 = fma(-d2, ys, as);
          const double num = fma(n1, D2, n2 * D1);
          acc[j] = fma(num, rcp_fast(D1 * D2), acc[j]);
        }
      }
    }
    const double scale = m.o2_abs_const * presda * (th * th * th);
    const double pdry_mb = pdrykpa * 10.0;
    const double n2base = m.n2_n * (m.n2_l * pdry_mb * pdry_mb * exp(m.n2_m * thln));
#pragma unroll
    for (int j = 0; j < FB; ++j) {
      const double f = fblk[j];
      const double f2 = f * f;
      const double nonres = m.o2_nonres * dfnr / (th * (f2 + dfnr * dfnr));
      double o2 = scale * (f2 * (nonres + acc[j]));
      o2 = (o2 < 0.0) ? 0.0 : o2;  // max(o2abs, 0) that keeps NaN (fmax would drop it)
      o2 *= m.o2_iso;
      double fdep = 1.0;
      if (m.n2_fdep) {
        const double fr = f / 450.0;
        fdep = 0.5 + 0.5 / (1.0 + fr * fr);
      }
      const double an2 = n2base * fdep * f2;
      if (f0 + j < F) adry[obase + (long long)j * L] = o2 + an2;
    }
  }

  // ------------------------------------------------------------------ cloud
  if (aliq) {
    const double denl = denliq ? denliq[ipt] : 0.0;
    if (denl > 0.0 || isnan(denl)) {
      if (m.liq_model == 0) {  // Liebe-91 double Debye
        const double theta1 = 1.0 - 300.0 / tk;
        const double eps0 = 77.66 - 103.3 * theta1;
        const double eps1 = 0.0671 * eps0;
        const double eps2 = 3.52;
        const double fp = (316.0 * theta1 + 146.4) * theta1 + 20.2;
        const double fs = 39.8 * fp;
#pragma unroll
        for (int j = 0; j < FB; ++j) {
          const double f = fblk[j];
          cplx eps = cadd(cdiv(cmake(eps0 - eps1, 0.0), cmake(1.0, f / fp)),
                          cdiv(cmake(eps1 - eps2, 0.0), cmake(1.0, f / fs)));
          eps.re += eps2;
          cplx re = cdiv(cmake(eps.re - 1.0, eps.im), cmake(eps.re + 2.0, eps.im));
          if (f0 + j < F) aliq[obase + (long long)j * L] = -0.06286 * re.im * f * denl;
        }
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
#pragma unroll 1
        for (int j = 0; j < FB; ++j) {
          const double f = fblk[j];
          const cplx z = cmake(0.0, f);
          cplx kappa = cmake(kap0, 0.0);
          kappa = csub(kappa, cdiv(cscale(z, delta), cmake(sd, f)));
          cplx chip = cdiv(cscale(clog_(cdiv(csub(z, z2), csub(z, z1))), hdelta), cnorm);
          cplx chij = cdiv(cscale(clog_(cdiv(csub(z, cconj(z2)), csub(z, cconj(z1)))), hdelta),
                           cconj(cnorm));
          cplx eps = cadd(kappa, cmake(chip.re + chij.re - deltb, chip.im + chij.im));
          cplx re = cdiv(cmake(eps.re - 1.0, eps.im), cmake(eps.re + 2.0, eps.im));
          if (f0 + j < F) aliq[obase + (long long)j * L] = -0.06286 * re.im * f * denl;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < FB; ++j)
        if (f0 + j < F) aliq[obase + (long long)j * L] = 0.0;
    }
  }
  if (aice) {
    const double deni = denice ? denice[ipt] : 0.0;
#pragma unroll
    for (int j = 0; j < FB; ++j) {
      const double f = fblk[j];
      const double wave = kLightC / (f * 1e9);
      double a = 0.0;
      if (deni > 0.0 || isnan(deni)) a = ((8.18645 / wave) * deni * 0.000959553) * kDb2Np;
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
  size_t smem = sizeof(double) * (size_t)(m.n_table + FB + 2 * m.n_h2o * FB);
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
