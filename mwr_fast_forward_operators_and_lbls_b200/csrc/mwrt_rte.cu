// mwrt_rte.cu -- kernel 3 with plain result stores (the template family lives in
// mwrt_rte_kernel.cuh), kernel 2 (slant path), the unfused band-pass fallback and the cloud mean
// radiating temperature (`tmrcld`).
#include "mwrt_rte_kernel.cuh"

namespace mwrt {

int rte_launch_fanout(const RteArgs& g, int from_sat, const Fanout& fo, cudaStream_t st);   // mwrt_rte_fanout.cu

// g.ds != null: general path (P,A,L).  g.ds == null: plane-parallel, z (P,L) + amass (A,) device arrays.
int rte_launch(const RteArgs& g, int from_sat, const Fanout& fo, cudaStream_t st) {
  if (g.bpC > 0 && (fo.mode != MWRT_OUT_PLAIN || g.ds != nullptr || g.bpCpb <= 0))
    return fail_arg(MWRT_E_STATE, "rte_launch: fused band-pass needs plain stores and the plane-parallel path");
  if (fo.mode != MWRT_OUT_PLAIN) return rte_launch_fanout(g, from_sat, fo, st);
  return rte_launch_t<false>(g, from_sat, fo, st);
}

// ------------------------------------------------------------------------------------------------
// Kernel 2 -- slant path (RTEquation.refractivity + RTEquation.ray_tracing; NOAA
// REFRACT / RAYTRAC), warp-cooperative:
//   CTA  = one profile; the refractive index n(level) and the layer term (n_{l-1} - n_l)/refbar_l do
//          not depend on the viewing angle and are computed once per profile into shared memory;
//   warp = one viewing angle, lane = level (rounds of 32 levels).  Every per-level quantity of the
//          ray trace depends on its own level and on the level below only, EXCEPT the bending angle
//          tau_l = sum_{k<=l} dtau_k, which is a warp-level inclusive prefix sum (5 shuffle steps per
//          round, the running total carried from round to round); "previous level" values come from
//          the neighbouring lane by shuffle; ds is stored coalesced along the level axis.
//   No trigonometric library call on the common path: the local elevation angle theta = 2 asin(s) is
//   never formed.  The branch test theta <= 2 theta0 is s <= sin(theta0); cos((theta+theta0)/4) and
//   cot(theta) follow from s by half-/double-angle identities (square roots only); the two sines of
//   tiny arguments and asin(sin(dtheta/4)) are short series (|x| < 0.05, else the library call).
//   Checked against the oracle's serial loop to 3e-14 relative (tests/test_kernel_algebra.py).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double refr_index(double p, double tk, double e) {
  const double pa = p - e;
  const double tc = tk - 273.16;
  const double tk2 = tk * tk;
  const double tc2 = tc * tc;
  const double rza = 1.0 + pa * (5.79e-07 * (1.0 + 0.52 / tk) - 0.00094611 * tc / tk2);
  const double rzw = 1.0 + 1650.0 * (e / (tk * tk2)) *
                               (1.0 - 0.01317 * tc + 0.000175 * tc2 + 1.44e-06 * (tc2 * tc));
  const double wetn = (64.79 * (e / tk) + 377600.0 * (e / tk2)) * rzw;
  const double dryn = 77.6036 * (pa / tk) * rza;
  return 1.0 + (dryn + wetn) * 1e-06;
}

// x + x^3/6 + 3x^5/40 + ... to x^11 (|x| < 0.05: next term < 2e-18 relative)
__device__ __forceinline__ double asin_small(double x) {
  const double x2 = x * x;
  double q = 945.0 / 42240.0;
  q = fma(q, x2, 105.0 / 3456.0);
  q = fma(q, x2, 15.0 / 336.0);
  q = fma(q, x2, 3.0 / 40.0);
  q = fma(q, x2, 1.0 / 6.0);
  return fma(x * x2, q, x);
}
// x - x^3/6 + ... to x^9 (|x| < 0.05)
__device__ __forceinline__ double sin_small(double x) {
  const double x2 = x * x;
  double q = 1.0 / 362880.0;
  q = fma(q, x2, -1.0 / 5040.0);
  q = fma(q, x2, 1.0 / 120.0);
  q = fma(q, x2, -1.0 / 6.0);
  return fma(x * x2, q, x);
}

constexpr int kSlantMaxWarps = 16;

__global__ void __launch_bounds__(32 * kSlantMaxWarps)
k_slant_path(const int P, const int L, const int A, const double* __restrict__ z,
             const double* __restrict__ p, const double* __restrict__ tk,
             const double* __restrict__ e, const double* __restrict__ angles,
             const int ray_tracing, double* __restrict__ ds, int* __restrict__ status) {
  extern __shared__ __align__(16) double sm[];
  double* zr = sm;            // L: height above the antenna
  double* nn = sm + L;        // L: refractive index
  double* dnr = sm + 2 * L;   // L: (n_{l-1} - n_l) / refbar_l   (layer l, l >= 1)
  const long long ip = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const double* zp = z + ip * L;
  const double z0 = zp[0];
  const double pi = 3.141592653589793;
  int znan = 0, badn = 0;
  for (int l = threadIdx.x; l < L; l += blockDim.x) {
    const double zi = zp[l] - z0;
    zr[l] = zi;
    znan |= isnan(zi);
    if (ray_tracing) {
      const double n = refr_index(p[ip * L + l], tk[ip * L + l], e[ip * L + l]);
      nn[l] = n;
      badn |= (n < 1.0);   // refractive index below 1 anywhere: PyRTlib warns and returns no path
    }
  }
  znan = __syncthreads_or(znan);
  badn = __syncthreads_or(badn);
  if (threadIdx.x == 0 && status) {
    if (znan) atomicOr(&status[ip], MWRT_ST_NAN_INPUT);
    if (badn) atomicOr(&status[ip], MWRT_ST_DUCTING);
  }
  if (ray_tracing && !badn) {
    for (int l = 1 + threadIdx.x; l < L; l += blockDim.x) {
      const double ni = nn[l], nl = nn[l - 1];
      double refbar;
      if (ni == nl || ni == 1.0 || nl == 1.0)
        refbar = (ni + nl) * 0.5;
      else
        refbar = 1.0 + (nl - ni) / (log((nl - 1.0) / (ni - 1.0)));
      dnr[l] = (nl - ni) / refbar;
    }
    __syncthreads();
  }
  const double re = kEarthRadius;
  for (int ia = warp; ia < A; ia += nw) {
    double* dsp = ds + (ip * A + ia) * (long long)L;
    const double angle = angles[ia];
    if (lane == 0) dsp[0] = 0.0;
    const bool vertical = (angle >= 89.0 && angle <= 91.0) || angle == -90.0;
    if (!ray_tracing || badn || vertical) {
      // plane-parallel diff(z)/sin(elev) (the reference's own calls); no path at all when n < 1
      // somewhere; the plain layer thickness within a degree of the zenith (as upstream)
      double scale = 1.0;
      if (!ray_tracing) scale = 1.0 / sin(angle * pi / 180.0);
      else if (badn) scale = 0.0;
      for (int l = 1 + lane; l < L; l += 32) dsp[l] = badn && ray_tracing ? 0.0 : (zr[l] - zr[l - 1]) * scale;
      continue;
    }
    const double theta0 = angle * (pi / 180.0);
    const double rs = re + zr[0] + z0;
    const double irs = 1.0 / rs;
    const double cos0 = cos(theta0), sin0 = sin(theta0);
    const double sina = sin(theta0 * 0.5);
    const double a0 = 2.0 * sina * sina;
    const double c40 = cos(theta0 * 0.25), s40 = sin(theta0 * 0.25);
    const double n0 = nn[0];
    // carried across rounds: values at the last level of the previous round
    double cot_prev = 1.0 / tan(theta0), tau_prev = 0.0, phi_prev = 0.0;
    bool ducted = false;
    for (int base = 1; base < L; base += 32) {
      const int l = base + lane;
      const bool valid = l < L;
      const int lc = valid ? l : L - 1;
      const double zi = zr[lc], zl = zr[lc - 1];
      const double ni = nn[lc];
      const double r = re + zi + z0, rl = re + zl + z0;
      // (divisions and square roots through the MUFU-seeded helpers of mwrt_common.cuh: every operand
      // here is a normal positive number, no IEEE slow path is needed)
      const double argdth = zi * irs - (n0 - ni) * cos0 * rcp_fast(ni);
      const double argth = 0.5 * (a0 + argdth) * rcp_fast(r);
      const bool ok = argth > 0.0;                       // else: ducting
      const double sint = sqrt_f(ok ? r * argth : 0.25); // sin(theta / 2)
      const double cosal = sqrt_f(fmax(1.0 - sint * sint, 1e-300));   // cos(theta / 2)
      double dtheta, cot;
      if (sint <= sin0) {                                // theta <= 2 theta0
        const double ch = sqrt_f(0.5 * (1.0 + cosal));   // cos(theta / 4)
        const double sh = sint * rcp_fast(2.0 * ch);     // sin(theta / 4)
        const double dendth = 2.0 * (sint + sina) * (ch * c40 - sh * s40);   // .. cos((theta + theta0) / 4)
        const double sind4 = div_fast(0.5 * argdth - zi * argth, dendth);
        dtheta = 4.0 * (fabs(sind4) < 0.05 ? asin_small(sind4) : asin(sind4));
        // cot(theta0 + dtheta) from sin(dtheta / 4) by two double-angle steps
        const double c4 = sqrt_f(fmax(1.0 - sind4 * sind4, 1e-300));
        const double s2 = 2.0 * sind4 * c4, c2 = 1.0 - 2.0 * sind4 * sind4;
        const double s1 = 2.0 * s2 * c2, c1 = 1.0 - 2.0 * s2 * s2;
        cot = div_fast(cos0 * c1 - sin0 * s1, sin0 * c1 + cos0 * s1);
      } else {
        dtheta = 2.0 * asin(sint) - theta0;
        cot = (1.0 - 2.0 * sint * sint) / (2.0 * sint * cosal);
      }
      const unsigned badmask = __ballot_sync(0xffffffffu, valid && !ok);
      double cot_l = __shfl_up_sync(0xffffffffu, cot, 1);
      if (lane == 0) cot_l = cot_prev;
      const double dtau = valid ? (cot + cot_l) * 0.5 * dnr[lc] : 0.0;
      // bending angle: inclusive prefix sum over the levels of this round + the carried total
      double tau = dtau;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(0xffffffffu, tau, o);
        if (lane >= o) tau += t;
      }
      tau += tau_prev;
      const double phi = dtheta + tau;
      double phil = __shfl_up_sync(0xffffffffu, phi, 1);
      if (lane == 0) phil = phi_prev;
      const double x = (phi - phil) * 0.5;
      const double shx = fabs(x) < 0.05 ? sin_small(x) : sin(x);
      const double dz = zi - zl;
      double dsi = sqrt_f(fmax(dz * dz + 4.0 * r * rl * (shx * shx), 1e-300));
      if (dtau != 0.0) {   // dtaua / (2 sin(dtaua / 2)), dtaua = |dtau|
        const double y = fabs(dtau) * 0.5;
        dsi = dsi * (y < 0.05 ? div_fast(y, sin_small(y)) : y / sin(y));
      }
      const int first = badmask ? __ffs(badmask) - 1 : 32;
      if (ducted || lane >= first) dsi = 0.0;   // the rest of the path stays 0, like upstream
      if (valid) dsp[l] = dsi;
      ducted = ducted || (badmask != 0u);
      cot_prev = __shfl_sync(0xffffffffu, cot, 31);
      tau_prev = __shfl_sync(0xffffffffu, tau, 31);
      phi_prev = __shfl_sync(0xffffffffu, phi, 31);
    }
    if (ducted && lane == 0 && status) atomicOr(&status[ip], MWRT_ST_DUCTING);
  }
}

int slant_path_launch(int P, int L, int A, const double* z, const double* p, const double* tk,
                      const double* e, const double* angles, int ray_tracing, double* ds, int* status,
                      cudaStream_t st) {
  const int nw = A < kSlantMaxWarps ? A : kSlantMaxWarps;
  const size_t smem = sizeof(double) * 3 * (size_t)L;
  if (smem > 48 * 1024) {
    if (smem > 200 * 1024) return fail_arg(MWRT_E_ARG, "slant path: more than 8533 levels per profile");
    MWRT_CUDA_OK(cudaFuncSetAttribute(k_slant_path, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  k_slant_path<<<(unsigned)P, 32 * nw, smem, st>>>(P, L, A, z, p, tk, e, angles, ray_tracing, ds, status);
  MWRT_CUDA_OK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Channel band-pass average as a separate pass -- the FALLBACK (sub-bands wider than one CTA, ray
// tracing, fan-out stores) and the stage entry point; the fused path reduces inside kernel 3 with the
// same chain.  Fixed summation order -> deterministic.
// ------------------------------------------------------------------------------------------------
__global__ void k_bandpass(const long long n, const int C, const int S, const int A,
                           const double* __restrict__ in, const double* __restrict__ w,
                           double* __restrict__ out, const Fanout fo) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // (p, c, a)
  if (idx >= n) return;
  const int a = (int)(idx % A);
  const long long pc = idx / A;
  const int c = (int)(pc % C);
  const long long p = pc / C;
  const double* src = in + ((p * C + c) * (long long)S) * A + a;
  double acc = 0.0;
  for (int s = 0; s < S; ++s) acc = fma(w ? w[c * S + s] : 1.0 / S, src[(long long)s * A], acc);
  out_store(out + idx, acc, fo);
}

int bandpass_launch(int P, int C, int S, int A, const double* in, const double* w, double* out,
                    const Fanout& fo, cudaStream_t st) {
  const long long n = (long long)P * C * A;
  k_bandpass<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, C, S, A, in, w, out, fo);
  MWRT_CUDA_OK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Mean radiating temperature of the (first) cloud layer -- column `tmrcld` of upstream's DataFrame
// (RTEquation.cloud_radiating_temperature; the reference reads `tbtotal` only,
// PyRTlib_processing.py:127).  Downwelling only.  Needs the running atmosphere term and optical depth
// at two levels per profile, which kernel 3 does not keep: a plain serial walk per (profile,
// frequency, angle) up to the cloud top, launched only when the column is asked for.
// ------------------------------------------------------------------------------------------------
__global__ void k_tmrcld(const RteArgs g, const int* __restrict__ ibase, const int* __restrict__ itop,
                         double* __restrict__ out) {
  __shared__ double etab[kExpTabN];
  __shared__ __align__(16) double ltab[128];
  exp_tab_stage(etab);
  log_tab_stage(ltab);
  __syncthreads();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // (p, f, a)
  if (idx >= (long long)g.P * g.F * g.A) return;
  const int a = (int)(idx % g.A);
  const long long pf = idx / g.A;
  const int jf = (int)(pf % g.F);
  const long long p = pf / g.F;
  const int L = g.L;
  const int ib = ibase[p], it = itop[p];
  double res = 0.0;
  if (ib >= 0 && it >= ib && it < L) {
    const double fq = g.frq[jf];
    const double hvk = (fq * 1e9) * kPlanckH / kBoltzmannK;
    const long long row = pf * L;
    const double* tkp = g.tk + p * L;
    const bool arrays = g.aliq != nullptr;
    const bool cloudy = arrays || g.denliq != nullptr;
    const double icec = ice_coef(fq);
    auto cloud = [&](int lv, double& al, double& ai) {
      al = 0.0;
      ai = 0.0;
      if (!cloudy) return;
      if (arrays) {
        al = g.aliq[row + lv];
        ai = g.aice[row + lv];
      } else {
        const double dl = g.denliq[p * L + lv], di = g.denice[p * L + lv];
        if (dl > 0.0 || isnan(dl)) al = liq_abs_unit(g.liq_model, tkp[lv], fq, etab, ltab) * dl;
        if (di > 0.0 || isnan(di)) ai = ice_abs_c(icec, di);
      }
    };
    bool neg = false;
    double tauprof = 0.0, batm = 0.0, tau_b = 0.0, batm_b = 0.0;
    double w0 = g.awet[row], d0 = g.adry[row], l0, i0;
    cloud(0, l0, i0);
    double bprev = 1.0 / (exp(hvk / tkp[0]) - 1.0);
    for (int i = 1; i <= it; ++i) {
      const double w1 = g.awet[row + i], d1 = g.adry[row + i];
      double l1, i1;
      cloud(i, l1, i1);
      const double dsl = g.ds ? g.ds[(p * g.A + a) * (long long)L + i]
                              : (g.z[p * L + i] - g.z[p * L + i - 1]) * g.amass[a];
      const double tau = layer_mean(w0, w1, true, neg, ltab) * dsl + layer_mean(d0, d1, true, neg, ltab) * dsl +
                         layer_mean(i0, i1, false, neg, ltab) * dsl + layer_mean(l0, l1, false, neg, ltab) * dsl;
      const double bcur = 1.0 / (exp(hvk / tkp[i]) - 1.0);
      const double ex = exp(-tau);
      const double boftlay = (bprev + bcur * ex) / (1.0 + ex);
      batm += boftlay * exp(-tauprof) * (1.0 - ex);
      tauprof += tau;
      if (i == ib) { tau_b = tauprof; batm_b = batm; }
      w0 = w1; d0 = d1; l0 = l1; i0 = i1; bprev = bcur;
    }
    if (tau_b > kExpMax) {
      res = 0.0;
    } else {
      const double batmcld = batm - batm_b, taucld = tauprof - tau_b;
      double boftcld = batmcld * exp(tau_b);
      if (!(taucld > kExpMax)) boftcld = boftcld / (1.0 - exp(-taucld));
      res = (boftcld > 0.0 || isnan(boftcld)) ? hvk / log(1.0 + 1.0 / boftcld) : 0.0;
    }
    if (neg) res = nan("");
  }
  if (g.status && (g.status[p] & (MWRT_ST_NAN_INPUT | MWRT_ST_RANGE | MWRT_ST_NEG_ABS))) res = nan("");
  out[idx] = res;
}

int tmrcld_launch(const RteArgs& g, const int* ibase, const int* itop, double* out, cudaStream_t st) {
  const long long n = (long long)g.P * g.F * g.A;
  k_tmrcld<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(g, ibase, itop, out);
  MWRT_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace mwrt
