// mwrt_rte_fanout.cu -- kernel 3 with its result stores fanned out to other GPUs (NVLink multicast or
// per-peer stores, staged per warp): the FANOUT = true half of the template family in
// mwrt_rte_kernel.cuh, in its own translation unit so the two halves compile in parallel.
#include "mwrt_rte_kernel.cuh"

namespace mwrt {

int rte_launch_fanout(int P, int L, int F, int A, const double* tk, const double* awet, const double* adry,
                      const double* aliq, const double* aice, const double* ds, const double* z,
                      const double* amass, const double* frq, int from_sat, double emis, double* tbtotal,
                      double* tbatm, double* tmr, double* tauwet, double* taudry, double* tauliq,
                      double* tauice, double* tau_layer, int tl_angle, int* status, const Fanout& fo,
                      cudaStream_t st) {
  return rte_launch_t<true>(P, L, F, A, tk, awet, adry, aliq, aice, ds, z, amass, frq, from_sat, emis, tbtotal,
                            tbatm, tmr, tauwet, taudry, tauliq, tauice, tau_layer, tl_angle, status, fo, st);
}

}  // namespace mwrt
