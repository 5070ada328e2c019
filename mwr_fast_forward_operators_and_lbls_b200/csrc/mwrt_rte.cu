// mwrt_rte.cu -- kernel 3 with plain result stores (the template family lives in
// mwrt_rte_kernel.cuh), kernel 2 (slant path) and the band-pass epilogue.
#include "mwrt_rte_kernel.cuh"

namespace mwrt {

int rte_launch_fanout(int P, int L, int F, int A, const double* tk, const double* awet, const double* adry,
                      const double* aliq, const double* aice, const double* ds, const double* z,
                      const double* amass, const double* frq, int from_sat, double emis, double* tbtotal,
                      double* tbatm, double* tmr, double* tauwet, double* taudry, double* tauliq,
                      double* tauice, double* tau_layer, int tl_angle, int* status, const Fanout& fo,
                      cudaStream_t st);   // mwrt_rte_fanout.cu

// ds != null: general path (P,A,L).  ds == null: plane-parallel, z (P,L) + amass (A,) device arrays.
int rte_launch(int P, int L, int F, int A, const double* tk, const double* awet, const double* adry,
               const double* aliq, const double* aice, const double* ds, const double* z,
               const double* amass, const double* frq, int from_sat, double emis, double* tbtotal,
               double* tbatm, double* tmr, double* tauwet, double* taudry, double* tauliq,
               double* tauice, double* tau_layer, int tl_angle, int* status, const Fanout& fo,
               cudaStream_t st) {
  if (fo.mode != MWRT_OUT_PLAIN)
    return rte_launch_fanout(P, L, F, A, tk, awet, adry, aliq, aice, ds, z, amass, frq, from_sat, emis, tbtotal,
                             tbatm, tmr, tauwet, taudry, tauliq, tauice, tau_layer, tl_angle, status, fo, st);
  return rte_launch_t<false>(P, L, F, A, tk, awet, adry, aliq, aice, ds, z, amass, frq, from_sat, emis, tbtotal,
                             tbatm, tmr, tauwet, taudry, tauliq, tauice, tau_layer, tl_angle, status, fo, st);
}

// ------------------------------------------------------------------------------------------------
// Slant path: thread = (profile, angle)
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

__global__ void k_slant_path(const int P, const int L, const int A, const double* __restrict__ z,
                             const double* __restrict__ p, const double* __restrict__ tk,
                             const double* __restrict__ e, const double* __restrict__ angles,
                             const int ray_tracing, double* __restrict__ ds,
                             int* __restrict__ status) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)P * A) return;
  const long long ip = idx / A;
  const int ia = (int)(idx - ip * A);
  const double* zp = z + ip * L;
  double* dsp = ds + idx * (long long)L;
  const double angle = angles[ia];
  const double z0 = zp[0];
  const double pi = 3.141592653589793;
  bool znan = isnan(z0);
  dsp[0] = 0.0;
  if (!ray_tracing) {
    const double amass = 1.0 / sin(angle * pi / 180.0);
    double zl = zp[0] - z0;
    for (int i = 1; i < L; ++i) {
      const double zi = zp[i] - z0;
      znan |= isnan(zi);
      dsp[i] = (zi - zl) * amass;
      zl = zi;
    }
    if (znan && status && ia == 0) atomicOr(&status[ip], MWRT_ST_NAN_INPUT);
    return;
  }
  const double* pp = p + ip * L;
  const double* tp = tk + ip * L;
  const double* ep = e + ip * L;
  // refractive index below 1 anywhere: PyRTlib warns and returns no path
  bool badn = false;
  for (int i = 0; i < L; ++i) {
    znan |= isnan(zp[i]);
    badn |= (refr_index(pp[i], tp[i], ep[i]) < 1.0);
  }
  if (znan && status && ia == 0) atomicOr(&status[ip], MWRT_ST_NAN_INPUT);
  for (int i = 1; i < L; ++i) dsp[i] = 0.0;
  if (badn) {
    if (status) atomicOr(&status[ip], MWRT_ST_DUCTING);
    return;
  }
  if ((angle >= 89.0 && angle <= 91.0) || angle == -90.0) {
    double zl = zp[0] - z0;
    for (int i = 1; i < L; ++i) {
      const double zi = zp[i] - z0;
      dsp[i] = zi - zl;
      zl = zi;
    }
    return;
  }
  const double re = kEarthRadius;
  const double theta0 = angle * (pi / 180.0);
  const double zr0 = zp[0] - z0;  // 0 unless NaN
  const double rs = re + zr0 + z0;
  const double costh0 = cos(theta0);
  const double sina = sin(theta0 * 0.5);
  const double a0 = 2.0 * sina * sina;
  double phil = 0.0, taul = 0.0, rl = re + zr0 + z0, tanthl = tan(theta0);
  const double n0 = refr_index(pp[0], tp[0], ep[0]);
  double nl_ = n0, zl = zr0;
  for (int i = 1; i < L; ++i) {
    const double zi = zp[i] - z0;
    const double ni = refr_index(pp[i], tp[i], ep[i]);
    const double r = re + zi + z0;
    double refbar;
    if (ni == nl_ || ni == 1.0 || nl_ == 1.0)
      refbar = (ni + nl_) * 0.5;
    else
      refbar = 1.0 + (nl_ - ni) / (log((nl_ - 1.0) / (ni - 1.0)));
    const double argdth = zi / rs - (n0 - ni) * costh0 / ni;
    const double argth = 0.5 * (a0 + argdth) / r;
    if (argth <= 0.0) {
      if (status) atomicOr(&status[ip], MWRT_ST_DUCTING);
      return;  // ducting: the rest of ds stays 0, like upstream
    }
    const double sint = sqrt(r * argth);
    double theta = 2.0 * asin(sint);
    double dtheta;
    if ((theta - 2.0 * theta0) <= 0.0) {
      const double dendth = 2.0 * (sint + sina) * cos((theta + theta0) * 0.25);
      const double sind4 = (0.5 * argdth - zi * argth) / dendth;
      dtheta = 4.0 * asin(sind4);
      theta = theta0 + dtheta;
    } else {
      dtheta = theta - theta0;
    }
    const double tanth = tan(theta);
    const double cthbar = ((1.0 / tanth) + (1.0 / tanthl)) * 0.5;
    const double dtau = cthbar * (nl_ - ni) / refbar;
    const double tau = taul + dtau;
    const double phi = dtheta + tau;
    const double sh = sin((phi - phil) * 0.5);
    const double dz = zi - zl;
    double dsi = sqrt(dz * dz + 4.0 * r * rl * (sh * sh));
    if (dtau != 0.0) {
      const double dtaua = fabs(tau - taul);
      dsi = dsi * (dtaua / (2.0 * sin(dtaua * 0.5)));
    }
    dsp[i] = dsi;
    phil = phi; taul = tau; rl = r; tanthl = tanth; nl_ = ni; zl = zi;
  }
}

int slant_path_launch(int P, int L, int A, const double* z, const double* p, const double* tk,
                      const double* e, const double* angles, int ray_tracing, double* ds, int* status,
                      cudaStream_t st) {
  const long long n = (long long)P * A;
  k_slant_path<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(P, L, A, z, p, tk, e, angles, ray_tracing,
                                                           ds, status);
  MWRT_CUDA_OK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Channel band-pass average (BASELINE.json config 3): fixed summation order -> deterministic
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

}  // namespace mwrt
