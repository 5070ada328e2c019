// mwrt_rte_args.h -- the argument block of kernel 3 (k_rte), shared by the kernel's translation units
// and the C-ABI layer (mwrt_api.cu) that fills it.
#pragma once

namespace mwrt {

#ifndef MWRT_RTE_THREADS
#define MWRT_RTE_THREADS 128
#endif
constexpr int kRteThreads = MWRT_RTE_THREADS;
constexpr int kRteBpMaxThreads = 384;   // fused band-pass: one CTA = whole channels, <= this many sub-frequencies

// Everything kernel 3 is given, as one by-value argument (the launcher family passes it through).
struct RteArgs {
  int P, L, F, A;
  const double *tk, *awet, *adry;
  const double *aliq, *aice;        // (P,F,L) cloud absorption arrays (stage entry point), or null:
  const double *denliq, *denice;    // (P,L) liquid / ice density -> absorption evaluated inline
  int liq_model;
  const double *ds;                 // (P,A,L) slant path, or null = plane-parallel from z and amass
  const double *z, *amass, *frq;
  double emis;
  double *tbtotal, *tbatm, *tmr, *tauwet, *taudry, *tauliq, *tauice;   // (P,F,A) [(P,C,A) with band-pass]
  double *tau_layer;                // (P,F,L) or null
  int tl_angle;
  int *status;
  // fused channel band-pass (BP instantiations): C channels x S sub-frequencies, cpb whole channels per
  // CTA, weights (C,S) or null = 1/S
  int bpC, bpS, bpCpb;
  const double *bpw;
};

// cpb (whole channels per CTA) for the fused band-pass, or 0 when a sub-band does not fit one CTA
inline int rte_bp_channels_per_cta(int C, int S) {
  if (C <= 0 || S <= 0 || S > kRteBpMaxThreads) return 0;
  const int cpb = kRteBpMaxThreads / S;
  return cpb > C ? C : cpb;
}

}  // namespace mwrt
