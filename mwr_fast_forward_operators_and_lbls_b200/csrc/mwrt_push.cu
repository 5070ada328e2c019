// mwrt_push.cu -- the multi-GPU gather as a small PUSH kernel over NVLink.
//
// Every rank owns one slot of a symmetric (world, fields, P, F, A) buffer; kernel 3 writes the rank's
// results into its own slot, and this kernel then delivers the slot to all peers:
//   multicast  one 16-byte `multimem.st` per 16 bytes to the NVLink MULTICAST address of the buffer --
//              NVSwitch replicates the store into every rank's copy, so a rank sends its data ONCE
//              (56 MB per step for config 2) instead of once per peer;
//   peers      where multicast is unavailable: one plain 16-byte store per peer-mapped copy.
// The grid is a handful of CTAs (a rank needs ~50 GB/s of egress to hide the gather behind the next
// step's kernels) launched on a high-priority side stream: it shares the SMs with the two FP64-bound
// kernels of the next step instead of taking 16+ CTAs for a collective.  Measured alternatives on
// 8 x B200 (profiles/r02e_*): NCCL all-gather 1.34 ms per step, copy-engine pushes (one cudaMemcpyAsync
// per peer) 1.87 ms -- seven concurrent peer copies reach only ~210 GB/s per GPU -- against 1.14 ms of
// kernels.
#include "mwrt_common.cuh"

namespace mwrt {

__device__ __forceinline__ void st_mc16(void* p, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

constexpr int kPushThreads = 256, kPushUnroll = 4;

// n16 = number of 16-byte words; dst: multicast address (MC) or up to MWRT_MAX_PEERS peer addresses
template <bool MC>
__global__ void __launch_bounds__(kPushThreads)
k_push(const float4* __restrict__ src, const long long n16, float4* mc_dst, const Fanout peers) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (kPushUnroll - 1) * stride < n16; i += kPushUnroll * stride) {
    float4 v[kPushUnroll];
#pragma unroll
    for (int u = 0; u < kPushUnroll; ++u) v[u] = src[i + u * stride];      // loads first: all in flight
#pragma unroll
    for (int u = 0; u < kPushUnroll; ++u) {
      if (MC) {
        st_mc16(mc_dst + i + u * stride, v[u]);
      } else {
        for (int r = 0; r < peers.n; ++r)
          reinterpret_cast<float4*>(peers.off[r])[i + u * stride] = v[u];
      }
    }
  }
  for (; i < n16; i += stride) {
    const float4 v = src[i];
    if (MC) {
      st_mc16(mc_dst + i, v);
    } else {
      for (int r = 0; r < peers.n; ++r) reinterpret_cast<float4*>(peers.off[r])[i] = v;
    }
  }
  __threadfence_system();   // the stores are ordered before whatever the stream signals next
}

int push_launch(const void* src, long long bytes, void* mc_dst, const long long* peer_dst, int npeer,
                int nblocks, cudaStream_t st) {
  if ((bytes & 15) || (reinterpret_cast<unsigned long long>(src) & 15))
    return fail_arg(MWRT_E_ARG, "push: source and size must be 16-byte aligned");
  if (nblocks < 1) nblocks = 8;
  const long long n16 = bytes / 16;
  Fanout f{MWRT_OUT_PEERS, 0, {0}};
  if (mc_dst) {
    if (reinterpret_cast<unsigned long long>(mc_dst) & 15) return fail_arg(MWRT_E_ARG, "push: unaligned destination");
    k_push<true><<<nblocks, kPushThreads, 0, st>>>(static_cast<const float4*>(src), n16,
                                                   static_cast<float4*>(mc_dst), f);
  } else {
    if (npeer < 1 || npeer > MWRT_MAX_PEERS || !peer_dst)
      return fail_arg(MWRT_E_ARG, "push: 1..MWRT_MAX_PEERS peer addresses required");
    f.n = npeer;
    for (int r = 0; r < npeer; ++r) {
      if (peer_dst[r] & 15) return fail_arg(MWRT_E_ARG, "push: unaligned destination");
      f.off[r] = peer_dst[r];   // absolute addresses here
    }
    k_push<false><<<nblocks, kPushThreads, 0, st>>>(static_cast<const float4*>(src), n16, nullptr, f);
  }
  MWRT_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace mwrt
