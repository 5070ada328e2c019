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
