// mwrt_rte_fanout.cu -- kernel 3 with its result stores fanned out to other GPUs (NVLink multicast or
// per-peer stores, staged per warp): the FANOUT = true half of the template family in
// mwrt_rte_kernel.cuh, in its own translation unit so the two halves compile in parallel.
#include "mwrt_rte_kernel.cuh"

namespace mwrt {

int rte_launch_fanout(const RteArgs& g, int from_sat, const Fanout& fo, cudaStream_t st) {
  return rte_launch_t<true>(g, from_sat, fo, st);
}

}  // namespace mwrt
