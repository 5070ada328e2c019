// mwrt_post.cu -- the rows either side of the hot path: post-processing statistics.
//
// k_stats: model-minus-reference statistics per column, the arithmetic of `stats_by_channel`
//   (/root/reference/python_src/plot_scripts/multi_campaign_plots_and_ana.py:155-178):
//     n_valid = #non-NaN(values)            (the reference counts the MODEL values only)
//     bias = nansum(dev)/n_valid; std = sqrt(nansum((dev-bias)^2)/n_valid); rmse = sqrt(nansum(dev^2)/n_valid)
//   one CTA per column (channel x elevation x ...), fixed-order block reduction -> deterministic.
#include "mwrt_common.cuh"

namespace mwrt {

constexpr int kStatThreads = 128;

__device__ __forceinline__ double block_sum(double v, double* sh) {
  // fixed order: warp shuffle tree, then warp 0 adds the per-warp partials in index order
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0)
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i];
  if (threadIdx.x == 0) sh[0] = t;
  __syncthreads();
  t = sh[0];
  return t;
}

__global__ void __launch_bounds__(kStatThreads)
k_stats(const int T, const long long M, const double* __restrict__ values,
        const double* __restrict__ refs, double* __restrict__ bias, double* __restrict__ sd,
        double* __restrict__ rmse, double* __restrict__ nvalid) {
  __shared__ double sh[kStatThreads / 32];
  const long long col = blockIdx.x;
  if (col >= M) return;
  double s1 = 0.0, s2 = 0.0, n = 0.0;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const double v = values[(long long)t * M + col];
    const double d = v - refs[(long long)t * M + col];
    if (!isnan(v)) n += 1.0;
    if (!isnan(d)) {
      s1 += d;
      s2 = fma(d, d, s2);
    }
  }
  n = block_sum(n, sh);
  s1 = block_sum(s1, sh);
  s2 = block_sum(s2, sh);
  const double avg = s1 / n;
  double s3 = 0.0;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const double d = values[(long long)t * M + col] - refs[(long long)t * M + col];
    if (!isnan(d)) {
      const double c = d - avg;
      s3 = fma(c, c, s3);
    }
  }
  s3 = block_sum(s3, sh);
  if (threadIdx.x == 0) {
    bias[col] = avg;
    sd[col] = sqrt(s3 / n);
    rmse[col] = sqrt(s2 / n);
    if (nvalid) nvalid[col] = n;
  }
}

int stats_launch(int T, long long M, const double* values, const double* refs, double* bias,
                 double* sd, double* rmse, double* nvalid, cudaStream_t st) {
  k_stats<<<(unsigned)M, kStatThreads, 0, st>>>(T, M, values, refs, bias, sd, rmse, nvalid);
  MWRT_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace mwrt

// ------------------------------------------------------------------------------------------------
// N1: the numerical core of the reference's profile pre-processing, batched over
// sondes.  /root/reference/python_src/preproc/preprocessing4all.py:
//   k_thin_mean     running_mean_from_arrays :156-169 on the thinning pattern of :282-285
//   k_assemble      humidity conversions :104-130, jump check :315-331, add_clim2profiles :478-531,
//                   lowest2tenhPa :450-474, flip and lowest-n_levels selection :807-814
// ------------------------------------------------------------------------------------------------
namespace mwrt {

constexpr int kMaxThin = 256;   // thinned points per sonde (reference: <= 200)
constexpr int kMaxClim = 64;    // climatology levels (AFGL: 50)

__device__ __forceinline__ double seg_nanmean(const double* a, int lo, int hi) {
  double s = 0.0;
  int n = 0;
  for (int k = lo; k < hi; ++k) {
    const double v = a[k];
    if (!isnan(v)) { s += v; ++n; }
  }
  return n ? s / n : nan("");
}

__global__ void k_thin_mean(const int S, const int N, const int M, const double* __restrict__ h,
                            const double* __restrict__ t, const double* __restrict__ p,
                            const double* __restrict__ rh, const int* __restrict__ inds,
                            const int* __restrict__ n_inds, const double* __restrict__ z_offset,
                            double* __restrict__ zo, double* __restrict__ to, double* __restrict__ po,
                            double* __restrict__ ro) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)S * M) return;
  const int s = (int)(idx / M), i = (int)(idx - (long long)s * M);
  const int n = n_inds[s];
  if (i >= n) {
    zo[idx] = to[idx] = po[idx] = ro[idx] = nan("");
    return;
  }
  const int* id = inds + (long long)s * M;
  const long long base = (long long)s * N;
  const int ind = id[i];
  zo[idx] = h[base + ind] - (z_offset ? z_offset[s] : 0.0);
  if (i == 0 || i == n - 1) {
    to[idx] = t[base + ind];
    po[idx] = p[base + ind];
    ro[idx] = rh[base + ind];
  } else {
    const int lo = (ind + id[i - 1]) / 2, hi = (ind + id[i + 1]) / 2;   // int((a+b)/2), a,b >= 0
    to[idx] = seg_nanmean(t + base, lo, hi);
    po[idx] = seg_nanmean(p + base, lo, hi);
    ro[idx] = seg_nanmean(rh + base, lo, hi);
  }
}

__device__ __forceinline__ double cc_liq(double tc) {
  return 610.78 * exp(2.5e6 / 462 * (1 / 273.15 - 1 / (273.15 + tc)));
}

__global__ void k_assemble(const int S, const int M, const int K, const int n_levels,
                           const double min_p, const double* __restrict__ zt,
                           const double* __restrict__ tt, const double* __restrict__ pt,
                           const double* __restrict__ rt, const int* __restrict__ n_inds,
                           const double* __restrict__ cz_km, const double* __restrict__ cp,
                           const double* __restrict__ ct, const double* __restrict__ cgkg,
                           const double* __restrict__ crh, double* __restrict__ o_p,
                           double* __restrict__ o_t, double* __restrict__ o_ppmv,
                           double* __restrict__ o_h2o, double* __restrict__ o_z,
                           double* __restrict__ o_rh, int* __restrict__ ok) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  const int n = n_inds[s];
  const double* z = zt + (long long)s * M;
  const double* t = tt + (long long)s * M;
  const double* p = pt + (long long)s * M;
  const double* r = rt + (long long)s * M;
  const long long ob = (long long)s * n_levels;
  double P[kMaxThin + kMaxClim], T[kMaxThin + kMaxClim], Q[kMaxThin + kMaxClim],
      W[kMaxThin + kMaxClim], Z[kMaxThin + kMaxClim], R[kMaxThin + kMaxClim];
  // rh given as a fraction?  (np.all(rh <= 1.5) -> rh *= 100)
  bool frac = true;
  for (int i = 0; i < n; ++i) frac = frac && (r[i] <= 1.5);
  const double rs = frac ? 100.0 : 1.0;
  // jump check (profile rejected -> NaN)
  bool reject = false;
  for (int i = 1; i < n / 2; ++i)
    if (fabs(z[i] - z[i - 1]) > 500 || fabs(p[i] - p[i - 1]) > 50) reject = true;
  if (reject || n < 2) {
    for (int k = 0; k < n_levels; ++k)
      o_p[ob + k] = o_t[ob + k] = o_ppmv[ob + k] = o_h2o[ob + k] = o_z[ob + k] = o_rh[ob + k] = nan("");
    if (ok) ok[s] = 0;
    return;
  }
  // threshold level from p, water vapour and z
  int p_index = -1, z_index = -1, wv_index = 0;
  double pmin = 0.0, zmax = 0.0, wvmin = 0.0;
  bool have_wv = false;
  for (int i = 0; i < n; ++i) {
    if (!isnan(p[i]) && (p_index < 0 || p[i] < pmin)) { pmin = p[i]; p_index = i; }
    if (!isnan(z[i]) && (z_index < 0 || z[i] > zmax)) { zmax = z[i]; z_index = i; }
    const double e = cc_liq(t[i] - 273.15) * (r[i] * rs) / 100;
    const double ppmv = 1000000 * e / (p[i] * 100);
    if (!isnan(ppmv) && (!have_wv || ppmv < wvmin)) { wvmin = ppmv; have_wv = true; }
  }
  for (int i = 0; i < n; ++i) {
    const double e = cc_liq(t[i] - 273.15) * (r[i] * rs) / 100;
    if (1000000 * e / (p[i] * 100) <= 2 * wvmin) wv_index = i;
  }
  int thr_idx = p_index < wv_index ? p_index : wv_index;
  thr_idx = z_index < thr_idx ? z_index : thr_idx;
  double thr = p[thr_idx];
  if (thr < min_p) thr = min_p;
  if (thr > 200) thr = 200;
  // radiosonde part below the threshold, climatology above
  int m = 0;
  for (int i = 0; i < n; ++i)
    if (p[i] > thr) {
      const double rhv = r[i] * rs;
      const double e = cc_liq(t[i] - 273.15) * rhv / 100;
      const double q = (0.622 * e) / (p[i] * 100 - 0.3777 * e);
      P[m] = p[i]; T[m] = t[i]; W[m] = q / (1 - q); Z[m] = z[i]; R[m] = rhv;
      ++m;
    }
  for (int k = 0; k < K; ++k)
    if (thr > cp[k]) {
      P[m] = cp[k]; T[m] = ct[k]; W[m] = cgkg[k] / 1000; Z[m] = cz_km[k] * 1000; R[m] = crh[k];
      ++m;
    }
  for (int i = 0; i < m; ++i) Q[i] = 1000000 * (cc_liq(T[i] - 273.15) * R[i] / 100) / (P[i] * 100);
  // lowest2tenhPa: make sure the top of the kept levels is below 10 hPa
  if (m > 179 && P[179] > 10) {
    int idx10 = -1;
    for (int i = 0; i < m && idx10 < 0; ++i)
      if (P[i] < 10) idx10 = i;
    if (idx10 >= 0) {
      int tg[7];
      const double step = (idx10 - 173) / 6.0;
      for (int q = 0; q < 7; ++q) {
        double v = (q == 6) ? (double)idx10 : 173 + q * step;
        int ti = (int)rint(v);
        tg[q] = ti < 0 ? 0 : (ti > m - 1 ? m - 1 : ti);
      }
      double b[6][7];
      for (int q = 0; q < 7; ++q) {
        b[0][q] = P[tg[q]]; b[1][q] = T[tg[q]]; b[2][q] = Q[tg[q]];
        b[3][q] = W[tg[q]]; b[4][q] = Z[tg[q]]; b[5][q] = R[tg[q]];
      }
      for (int q = 0; q < 7; ++q) {
        P[173 + q] = b[0][q]; T[173 + q] = b[1][q]; Q[173 + q] = b[2][q];
        W[173 + q] = b[3][q]; Z[173 + q] = b[4][q]; R[173 + q] = b[5][q];
      }
    }
  }
  // flip (index 0 = top) and keep the lowest n_levels
  for (int k = 0; k < n_levels; ++k) {
    const int src = n_levels - 1 - k;   // ground-first index
    const bool have = src < m;
    o_p[ob + k] = have ? P[src] : nan("");
    o_t[ob + k] = have ? T[src] : nan("");
    o_ppmv[ob + k] = have ? Q[src] : nan("");
    o_h2o[ob + k] = have ? 1000 * W[src] : nan("");
    o_z[ob + k] = have ? Z[src] : nan("");
    o_rh[ob + k] = have ? R[src] : nan("");
  }
  if (ok) ok[s] = 1;
}

int thin_mean_launch(int S, int N, int M, const double* h, const double* t, const double* p,
                     const double* rh, const int* inds, const int* n_inds, const double* z_offset,
                     double* zo, double* to, double* po, double* ro, cudaStream_t st) {
  const long long n = (long long)S * M;
  k_thin_mean<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(S, N, M, h, t, p, rh, inds, n_inds, z_offset,
                                                          zo, to, po, ro);
  MWRT_CUDA_OK(cudaGetLastError());
  return 0;
}

int assemble_launch(int S, int M, int K, int n_levels, double min_p, const double* zt, const double* tt,
                    const double* pt, const double* rt, const int* n_inds, const double* cz_km,
                    const double* cp, const double* ct, const double* cgkg, const double* crh,
                    double* o_p, double* o_t, double* o_ppmv, double* o_h2o, double* o_z, double* o_rh,
                    int* ok, cudaStream_t st) {
  if (M > kMaxThin || K > kMaxClim) return fail_arg(MWRT_E_ARG, "assemble: M <= 256 and K <= 64 required");
  k_assemble<<<(S + 31) / 32, 32, 0, st>>>(S, M, K, n_levels, min_p, zt, tt, pt, rt, n_inds, cz_km, cp, ct,
                                           cgkg, crh, o_p, o_t, o_ppmv, o_h2o, o_z, o_rh, ok);
  MWRT_CUDA_OK(cudaGetLastError());
  return 0;
}


// ------------------------------------------------------------------------------------------------
// N3: batched cloud detection + adiabatic water content, thread = one profile.
// /root/reference/python_src/preproc/derive_cloud_water.py: derive_cloud_features :146-363 (relative-
// humidity layers after Nandan et al. 2022, with the reference's own list handling) and calc_lwc
// :68-142.  Arrays are TOP-first (index 0 = top), like the reference's.  The reference's `cloud_bools`
// is bookkeeping no output depends on and is not kept.
// ------------------------------------------------------------------------------------------------
constexpr int kCloudMaxLevels = 256;
constexpr int kCloudMaxLayers = 128;

__device__ __forceinline__ int cloud_band(double z) {   // strict limits: -1 exactly on a band edge
  if (z < 2000.0) return 0;
  if (2000.0 < z && z < 6000.0) return 1;
  if (6000.0 < z && z < 12000.0) return 2;
  if (12000.0 < z) return 3;
  return -1;
}
__device__ __forceinline__ int cloud_nearest(const double* z, int L, double x) {   // np.nanargmin(|z - x|)
  int best = 0;
  double bv = nan("");
  for (int i = 0; i < L; ++i) {
    const double d = fabs(z[i] - x);
    if (!isnan(d) && (isnan(bv) || d < bv)) { bv = d; best = i; }
  }
  return best;
}

__global__ void k_cloud_features(const int S, const int L, const double* __restrict__ p_in,
                                 const double* __restrict__ t_in, const double* __restrict__ z_in,
                                 const double* __restrict__ rh_in, double* __restrict__ lwc_m3,
                                 double* __restrict__ lwc_kg, double* __restrict__ iwc_m3,
                                 double* __restrict__ iwc_kg, double* __restrict__ lwp,
                                 double* __restrict__ iwp) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  const double* p = p_in + (long long)s * L;
  const double* t = t_in + (long long)s * L;
  const double* z = z_in + (long long)s * L;
  const long long ob = (long long)s * L;
  const double rh_min[4] = {92, 90, 88, 75}, rh_max[4] = {95, 93, 90, 80}, rh_int[4] = {84, 82, 78, 70};
  double rh[kCloudMaxLevels];
  double tops[kCloudMaxLayers], bases[kCloudMaxLayers];
  bool rmb[kCloudMaxLayers], rmt[kCloudMaxLayers];
  int n = 0, nb = 0;   // tops, bases
  // 1) RH with respect to ice below 0 C
  for (int i = 0; i < L; ++i) {
    double r = rh_in[ob + i];
    const double tk = t[i];
    if (tk < 273.15) {
      const double tc = tk - 273.15;
      const double el = 610.78 * exp(2.5e6 / 462 * (1 / 273.15 - 1 / (273.15 + tc)));
      const double ei = 610.78 * exp(2.840e6 / 462 * (1 / 273.15 - 1 / (273.15 + tc)));
      r = r * el / ei;
    }
    rh[i] = r;
    lwc_m3[ob + i] = 0.0; lwc_kg[ob + i] = 0.0; iwc_m3[ob + i] = 0.0; iwc_kg[ob + i] = 0.0;
  }
  // 2-4) preliminary layers: RH above RH_min of the altitude band
  bool in_cloud = false;
  for (int i = 0; i < L; ++i) {
    const int b = cloud_band(z[i]);
    if (b < 0) continue;
    if (rh[i] > rh_min[b]) {
      if (!in_cloud && n < kCloudMaxLayers) tops[n++] = z[i];
      in_cloud = true;
    } else {
      if (in_cloud && nb < kCloudMaxLayers) bases[nb++] = z[i];
      in_cloud = false;
    }
  }
  if (n == nb + 1 && nb < kCloudMaxLayers) bases[nb++] = z[L - 1];
  n = n < nb ? n : nb;
  // 5) bases below 500 m with a thickness under 400 m
  int m = 0;
  for (int i = 0; i < n; ++i)
    if (!(bases[i] < 500.0 && fabs(bases[i] - tops[i]) < 400.0)) { bases[m] = bases[i]; tops[m] = tops[i]; ++m; }
  n = m;
  // 6) RH_max must be reached inside the layer
  m = 0;
  for (int i = 0; i < n; ++i) {
    bool keep = true;
    const int b = cloud_band(bases[i]);
    if (b >= 0) {
      const int it = cloud_nearest(z, L, tops[i]), ib = cloud_nearest(z, L, bases[i]);
      bool any = false;
      for (int k = it; k < ib; ++k) any = any || (rh_max[b] < rh[k]);
      keep = any;
    }
    if (keep) { bases[m] = bases[i]; tops[m] = tops[i]; ++m; }
  }
  n = m;
  // 7) connect layers: adjacent levels, a gap under 300 m, or a gap that stays above RH_inter
  for (int i = 0; i < n; ++i) { rmb[i] = false; rmt[i] = false; }
  double rh_inter = nan("");
  for (int i = 0; i < n; ++i) {
    const int b = cloud_band(bases[i]);
    if (b >= 0) rh_inter = rh_int[b];
    if (i != 0) {
      const int ib = cloud_nearest(z, L, bases[i - 1]), it = cloud_nearest(z, L, tops[i]);
      bool merge = false;
      if (abs(ib - it) == 1) {
        merge = rh[ib] > rh_inter;
      } else if (fabs(bases[i - 1] - tops[i]) < 300.0) {
        merge = true;
      } else if (it > ib) {
        double mn = nan("");
        for (int k = ib; k < it; ++k)
          if (!isnan(rh[k]) && (isnan(mn) || rh[k] < mn)) mn = rh[k];
        merge = mn > rh_inter;
      }
      if (merge) { rmb[i - 1] = true; rmt[i] = true; }
    }
  }
  {
    int mb = 0, mt = 0;
    for (int i = 0; i < n; ++i) {
      if (!rmb[i]) bases[mb++] = bases[i];
      if (!rmt[i]) tops[mt++] = tops[i];
    }
    n = mb < mt ? mb : mt;
  }
  // 8) thinner than 100 m
  m = 0;
  for (int i = 0; i < n; ++i)
    if (!(fabs(bases[i] - tops[i]) < 100.0)) { bases[m] = bases[i]; tops[m] = tops[i]; ++m; }
  n = m;
  // calc_lwc (Chakraborty & Maitra 2011 / Nandan et al. 2022)
  const double cp = 1003.5, Lh = 334944.0, R_L = 287.06, dgam = 9.76e-3 - 6.5e-3;
  for (int i = 0; i < n; ++i) {
    const int it = cloud_nearest(z, L, tops[i]), ib = cloud_nearest(z, L, bases[i]);
    double* dst_m3;
    double* dst_kg;
    if (t[ib] > 273.15 && t[it] > 273.15) { dst_m3 = lwc_m3; dst_kg = lwc_kg; }
    else if (t[ib] < 233.15 && t[it] < 233.15) { dst_m3 = iwc_m3; dst_kg = iwc_kg; }
    else if (t[ib] > 233.15 && t[it] < 273.15) { dst_m3 = lwc_m3; dst_kg = lwc_kg; }
    else continue;
    for (int j = it; j < ib; ++j) {
      const double rho = p[j] * 100 / R_L / t[j];
      const double dz = z[j == 0 ? L - 1 : j - 1] - z[j];   // the reference's z_array[j-1] wraps at j = 0
      const double ad = rho * cp / Lh * dgam * dz;
      const double dh = z[j] - bases[i];
      double w = ad * (1.239 - 0.145 * log(dh));
      if (w < 0) w = 0;
      dst_m3[ob + j] = w;
      dst_kg[ob + j] = w / rho;
    }
  }
  double sl = 0.0, si = 0.0;
  for (int i = 0; i < L; ++i) {   // np.gradient(z): central differences, one-sided at the ends
    double g;
    if (L == 1) g = 0.0;
    else if (i == 0) g = z[1] - z[0];
    else if (i == L - 1) g = z[L - 1] - z[L - 2];
    else g = (z[i + 1] - z[i - 1]) / 2.0;
    sl += lwc_m3[ob + i] * g;
    si += iwc_m3[ob + i] * g;
  }
  if (lwp) lwp[s] = fabs(sl);
  if (iwp) iwp[s] = fabs(si);
}

int cloud_features_launch(int S, int L, const double* p, const double* t, const double* z, const double* rh,
                          double* lwc_m3, double* lwc_kg, double* iwc_m3, double* iwc_kg, double* lwp,
                          double* iwp, cudaStream_t st) {
  if (L > kCloudMaxLevels) return fail_arg(MWRT_E_ARG, "cloud features: at most 256 levels");
  k_cloud_features<<<(S + 31) / 32, 32, 0, st>>>(S, L, p, t, z, rh, lwc_m3, lwc_kg, iwc_m3, iwc_kg, lwp, iwp);
  MWRT_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace mwrt
