// mwrt_common.cuh -- shared device helpers for libmwrt (sm_100a only).
//
// Everything here is fp64: the path is bound by the FP64 pipe (64 lanes/clk/SM on B200) and MUFU,
// not by HBM or tensor cores (DESIGN.md "Roofline").  The helpers trade IEEE corner cases the
// physics cannot reach (denormals, infinities in denominators) for fewer FP64 issue slots.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "../../include/mwrt.h"

namespace mwrt {

// physical constants -- the values PyRTlib's utils.constants() carries (oracle/pyrtlib_scalar.py)
constexpr double kPlanckH = 6.6260755e-34;
constexpr double kBoltzmannK = 1.380658e-23;
constexpr double kLightC = 299792458.0;
constexpr double kTCosmic = 2.736;   // constants('Tcosmicbkg'): 2.736 +/- 0.017 K (Janssen 1993, p. 12)
constexpr double kEarthRadius = 6370.949;
constexpr double kRWatVap = 461.5;
constexpr double kSqrtPi = 1.77245385090551603;
constexpr double kExpMax = 125.0;
constexpr double kDb2Np = 0.2302585092994046;  // ln(10)/10
constexpr double kRvapAbs = (0.01 * 8.314510) / 18.01528;

// Device-side view of a model: scalars by value, line tables by device pointer (column-major,
// MWRT_H2O_NCOL x n_h2o then MWRT_O2_NCOL x n_o2 in ONE allocation so a CTA stages it with one loop).
struct ModelDev {
  int n_h2o, n_o2;
  int h2o_sd, o2_mixing, o2_dens118, o2_sd118, n2_fdep, liq_model;
  double h2o_reftline, h2o_reftcon, h2o_cf, h2o_xcf, h2o_cs, h2o_xcs, h2o_pvap_div, h2o_den_const,
      h2o_line_const;
  double o2_x, o2_wb300, o2_pvap_div, o2_h2o_broad, o2_nonres, o2_abs_const, o2_iso;
  double n2_l, n2_m, n2_n;
  const double* tables;  // device
  int n_table;           // doubles in `tables`
  // the same tables LINE-major, as kernel 1 stages them: n_h2o rows of kHS doubles (the columns, then
  // 1/fl^2, padding) followed by the O2 rows of kOS doubles (columns, 1/f^2, padding), O2 padded to an
  // even line count with a zero-strength copy of the last line.  Built once at model creation.
  const double* rows;    // device
  int n_rows;            // doubles in `rows`
};
// Per-plan cache of kernel 1's per-frequency-block scratch (k_abs_blocks in mwrt_absorption.cu): device
// buffer + what it currently holds.
struct AbsBlocks {
  double* buf = nullptr;
  size_t cap = 0;                          // doubles
  int fb = 0, f32 = -1, far = -1, F = 0;
  const double* frq = nullptr;
};
constexpr int kHS = MWRT_H2O_NCOL + 2;   // 20
constexpr int kOS = MWRT_O2_NCOL + 2;    // 12
constexpr int kH_IFL2 = MWRT_H2O_NCOL, kO_IFK2 = MWRT_O2_NCOL;

// ---- result stores fused with the multi-GPU gather ------------------------------------------------
// mode 0: plain store.  mode 1: the output pointer is an NVLink MULTICAST address (NVSwitch replicates
// one `multimem.st` into the same offset of every rank's symmetric buffer).  mode 2: one plain store
// per peer, at the pointer plus that peer's byte offset (peer-mapped symmetric buffers; offset 0 = own).
// Either way kernel 3 itself delivers its results to every GPU and no separate collective runs.
struct Fanout {
  int mode, n;
  long long off[MWRT_MAX_PEERS];
};
__device__ __forceinline__ void out_store_mc(double* p, double v) {
  asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
// 128-bit multicast store (16-byte aligned); multimem.st has no .v2.f64 form, the bits travel as four
// .f32 lanes
__device__ __forceinline__ void out_store2_mc(double* p, double v0, double v1) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p),
               "f"(__int_as_float(__double2loint(v0))), "f"(__int_as_float(__double2hiint(v0))),
               "f"(__int_as_float(__double2loint(v1))), "f"(__int_as_float(__double2hiint(v1)))
               : "memory");
}
__device__ __forceinline__ void out_store(double* p, double v, const Fanout& fo) {
  if (fo.mode == MWRT_OUT_PLAIN) {
    *p = v;
  } else if (fo.mode == MWRT_OUT_MULTICAST) {
    out_store_mc(p, v);
  } else {
#pragma unroll 1
    for (int r = 0; r < fo.n; ++r) *reinterpret_cast<double*>(reinterpret_cast<char*>(p) + fo.off[r]) = v;
  }
}

// ---- reciprocal / division without the IEEE slow path --------------------------------------
// MUFU.RCP64H seed (~20 bits) + one cubic Newton step (3 DFMA) -> ~2^-60 before rounding.
// Valid for normal, finite x (all line-shape denominators are >= width^2 > 0).
__device__ __forceinline__ double rcp_fast(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  double t = fma(e, e, e);
  return fma(r, t, r);
}
// Same seed, ONE quadratic Newton step (2 DFMA): relative error e^2 where e is the seed's error.
// MUFU.RCP64H is good to 2^-20 (measured worst case 2^-19.97, tests/test_gpu_parity.py), so this is
// accurate to 1e-12 worst case (~3e-13 rms) -- three orders inside the 1e-9 parity budget.  Used only in the
// two innermost loops (line-pair sums, per-angle RTE stage), where it removes one FP64 instruction
// per reciprocal.
__device__ __forceinline__ double rcp_seed(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  return r;
}
__device__ __forceinline__ double rcp_nr1(double x) {
  const double r = rcp_seed(x);
  return fma(r, fma(-x, r, 1.0), r);
}
// a/b with one residual correction (result within 1 ulp of IEEE)
__device__ __forceinline__ double div_fast(double a, double b) {
  double r = rcp_fast(b);
  double q = a * r;
  double rem = fma(-b, q, a);
  return fma(rem, r, q);
}

// ---- ordered FP64 ops for the staged helpers (exp_tb_vec, rcp_nr1_vec) ---------------------------------
// Those evaluate N independent arguments stage by stage, so that the N dependency chains are interleaved
// in program order (the FP64 pipe needs ~4 independent DFMAs in flight per scheduler).  `asm volatile`
// keeps the compiler front end from re-serialising the chains (left alone it schedules depth-first to
// save registers; measured: 52 % of FP64 instructions issued back-to-back dependent).  ptxas still owns
// the final schedule but starts from this order.
#define MWRT_FMA(d, a, b, c) asm volatile("fma.rn.f64 %0, %1, %2, %3;" : "=d"(d) : "d"(a), "d"(b), "d"(c))
#define MWRT_MUL(d, a, b) asm volatile("mul.rn.f64 %0, %1, %2;" : "=d"(d) : "d"(a), "d"(b))
#define MWRT_ADD(d, a, b) asm volatile("add.rn.f64 %0, %1, %2;" : "=d"(d) : "d"(a), "d"(b))

// ---- table-driven exp: 8 FP64-pipe instructions -------------------------------------------------
// exp(x) = 2^k * T[j] * exp(r),  n = round(x*256/ln2) = 256 k + j,  r = x - n*ln2/256, |r| <= ln2/512:
// exp(r)-1 needs only degree 4 (r^5/120 < 4e-17), T[j] = 2^(j/256) correctly rounded, read from a
// 256-entry shared-memory table (one LDS.64; the LSU pipe is idle in both hot kernels).  The
// reduction uses ONE constant, so the result carries a relative error of about |x| * 1.1e-16 on top
// of ~1 ulp -- the arguments here are optical depths and line-strength exponents of order 1..10.
// Domain: |x| < 700 (callers clamp); no NaN/inf/denormal handling.
constexpr int kExpTabN = 256;
static __device__ const double kExpTab[kExpTabN] = {
    0x1.0000000000000p+0, 0x1.00b1afa5abcbfp+0, 0x1.0163da9fb3335p+0, 0x1.02168143b0281p+0,
    0x1.02c9a3e778061p+0, 0x1.037d42e11bbccp+0, 0x1.04315e86e7f85p+0, 0x1.04e5f72f654b1p+0,
    0x1.059b0d3158574p+0, 0x1.0650a0e3c1f89p+0, 0x1.0706b29ddf6dep+0, 0x1.07bd42b72a836p+0,
    0x1.0874518759bc8p+0, 0x1.092bdf66607e0p+0, 0x1.09e3ecac6f383p+0, 0x1.0a9c79b1f3919p+0,
    0x1.0b5586cf9890fp+0, 0x1.0c0f145e46c85p+0, 0x1.0cc922b7247f7p+0, 0x1.0d83b23395decp+0,
    0x1.0e3ec32d3d1a2p+0, 0x1.0efa55fdfa9c5p+0, 0x1.0fb66affed31bp+0, 0x1.1073028d7233ep+0,
    0x1.11301d0125b51p+0, 0x1.11edbab5e2ab6p+0, 0x1.12abdc06c31ccp+0, 0x1.136a814f204abp+0,
    0x1.1429aaea92de0p+0, 0x1.14e95934f312ep+0, 0x1.15a98c8a58e51p+0, 0x1.166a45471c3c2p+0,
    0x1.172b83c7d517bp+0, 0x1.17ed48695bbc0p+0, 0x1.18af9388c8deap+0, 0x1.1972658375d2fp+0,
    0x1.1a35beb6fcb75p+0, 0x1.1af99f8138a1cp+0, 0x1.1bbe084045cd4p+0, 0x1.1c82f95281c6bp+0,
    0x1.1d4873168b9aap+0, 0x1.1e0e75eb44027p+0, 0x1.1ed5022fcd91dp+0, 0x1.1f9c18438ce4dp+0,
    0x1.2063b88628cd6p+0, 0x1.212be3578a819p+0, 0x1.21f49917ddc96p+0, 0x1.22bdda27912d1p+0,
    0x1.2387a6e756238p+0, 0x1.2451ffb82140ap+0, 0x1.251ce4fb2a63fp+0, 0x1.25e85711ece75p+0,
    0x1.26b4565e27cddp+0, 0x1.2780e341ddf29p+0, 0x1.284dfe1f56381p+0, 0x1.291ba7591bb70p+0,
    0x1.29e9df51fdee1p+0, 0x1.2ab8a66d10f13p+0, 0x1.2b87fd0dad990p+0, 0x1.2c57e39771b2fp+0,
    0x1.2d285a6e4030bp+0, 0x1.2df961f641589p+0, 0x1.2ecafa93e2f56p+0, 0x1.2f9d24abd886bp+0,
    0x1.306fe0a31b715p+0, 0x1.31432edeeb2fdp+0, 0x1.32170fc4cd831p+0, 0x1.32eb83ba8ea32p+0,
    0x1.33c08b26416ffp+0, 0x1.3496266e3fa2dp+0, 0x1.356c55f929ff1p+0, 0x1.36431a2de883bp+0,
    0x1.371a7373aa9cbp+0, 0x1.37f26231e754ap+0, 0x1.38cae6d05d866p+0, 0x1.39a401b7140efp+0,
    0x1.3a7db34e59ff7p+0, 0x1.3b57fbfec6cf4p+0, 0x1.3c32dc313a8e5p+0, 0x1.3d0e544ede173p+0,
    0x1.3dea64c123422p+0, 0x1.3ec70df1c5175p+0, 0x1.3fa4504ac801cp+0, 0x1.40822c367a024p+0,
    0x1.4160a21f72e2ap+0, 0x1.423fb2709468ap+0, 0x1.431f5d950a897p+0, 0x1.43ffa3f84b9d4p+0,
    0x1.44e086061892dp+0, 0x1.45c2042a7d232p+0, 0x1.46a41ed1d0057p+0, 0x1.4786d668b3237p+0,
    0x1.486a2b5c13cd0p+0, 0x1.494e1e192aed2p+0, 0x1.4a32af0d7d3dep+0, 0x1.4b17dea6db7d7p+0,
    0x1.4bfdad5362a27p+0, 0x1.4ce41b817c114p+0, 0x1.4dcb299fddd0dp+0, 0x1.4eb2d81d8abffp+0,
    0x1.4f9b2769d2ca7p+0, 0x1.508417f4531eep+0, 0x1.516daa2cf6642p+0, 0x1.5257de83f4eefp+0,
    0x1.5342b569d4f82p+0, 0x1.542e2f4f6ad27p+0, 0x1.551a4ca5d920fp+0, 0x1.56070dde910d2p+0,
    0x1.56f4736b527dap+0, 0x1.57e27dbe2c4cfp+0, 0x1.58d12d497c7fdp+0, 0x1.59c0827ff07ccp+0,
    0x1.5ab07dd485429p+0, 0x1.5ba11fba87a03p+0, 0x1.5c9268a5946b7p+0, 0x1.5d84590998b93p+0,
    0x1.5e76f15ad2148p+0, 0x1.5f6a320dceb71p+0, 0x1.605e1b976dc09p+0, 0x1.6152ae6cdf6f4p+0,
    0x1.6247eb03a5585p+0, 0x1.633dd1d1929fdp+0, 0x1.6434634ccc320p+0, 0x1.652b9febc8fb7p+0,
    0x1.6623882552225p+0, 0x1.671c1c70833f6p+0, 0x1.68155d44ca973p+0, 0x1.690f4b19e9538p+0,
    0x1.6a09e667f3bcdp+0, 0x1.6b052fa75173ep+0, 0x1.6c012750bdabfp+0, 0x1.6cfdcddd47645p+0,
    0x1.6dfb23c651a2fp+0, 0x1.6ef9298593ae5p+0, 0x1.6ff7df9519484p+0, 0x1.70f7466f42e87p+0,
    0x1.71f75e8ec5f74p+0, 0x1.72f8286ead08ap+0, 0x1.73f9a48a58174p+0, 0x1.74fbd35d7cbfdp+0,
    0x1.75feb564267c9p+0, 0x1.77024b1ab6e09p+0, 0x1.780694fde5d3fp+0, 0x1.790b938ac1cf6p+0,
    0x1.7a11473eb0187p+0, 0x1.7b17b0976cfdbp+0, 0x1.7c1ed0130c132p+0, 0x1.7d26a62ff86f0p+0,
    0x1.7e2f336cf4e62p+0, 0x1.7f3878491c491p+0, 0x1.80427543e1a12p+0, 0x1.814d2add106d9p+0,
    0x1.82589994cce13p+0, 0x1.8364c1eb941f7p+0, 0x1.8471a4623c7adp+0, 0x1.857f4179f5b21p+0,
    0x1.868d99b4492edp+0, 0x1.879cad931a436p+0, 0x1.88ac7d98a6699p+0, 0x1.89bd0a478580fp+0,
    0x1.8ace5422aa0dbp+0, 0x1.8be05bad61778p+0, 0x1.8cf3216b5448cp+0, 0x1.8e06a5e0866d9p+0,
    0x1.8f1ae99157736p+0, 0x1.902fed0282c8ap+0, 0x1.9145b0b91ffc6p+0, 0x1.925c353aa2fe2p+0,
    0x1.93737b0cdc5e5p+0, 0x1.948b82b5f98e5p+0, 0x1.95a44cbc8520fp+0, 0x1.96bdd9a7670b3p+0,
    0x1.97d829fde4e50p+0, 0x1.98f33e47a22a2p+0, 0x1.9a0f170ca07bap+0, 0x1.9b2bb4d53fe0dp+0,
    0x1.9c49182a3f090p+0, 0x1.9d674194bb8d5p+0, 0x1.9e86319e32323p+0, 0x1.9fa5e8d07f29ep+0,
    0x1.a0c667b5de565p+0, 0x1.a1e7aed8eb8bbp+0, 0x1.a309bec4a2d33p+0, 0x1.a42c980460ad8p+0,
    0x1.a5503b23e255dp+0, 0x1.a674a8af46052p+0, 0x1.a799e1330b358p+0, 0x1.a8bfe53c12e59p+0,
    0x1.a9e6b5579fdbfp+0, 0x1.ab0e521356ebap+0, 0x1.ac36bbfd3f37ap+0, 0x1.ad5ff3a3c2774p+0,
    0x1.ae89f995ad3adp+0, 0x1.afb4ce622f2ffp+0, 0x1.b0e07298db666p+0, 0x1.b20ce6c9a8952p+0,
    0x1.b33a2b84f15fbp+0, 0x1.b468415b749b1p+0, 0x1.b59728de5593ap+0, 0x1.b6c6e29f1c52ap+0,
    0x1.b7f76f2fb5e47p+0, 0x1.b928cf22749e4p+0, 0x1.ba5b030a1064ap+0, 0x1.bb8e0b79a6f1fp+0,
    0x1.bcc1e904bc1d2p+0, 0x1.bdf69c3f3a207p+0, 0x1.bf2c25bd71e09p+0, 0x1.c06286141b33dp+0,
    0x1.c199bdd85529cp+0, 0x1.c2d1cd9fa652cp+0, 0x1.c40ab5fffd07ap+0, 0x1.c544778fafb22p+0,
    0x1.c67f12e57d14bp+0, 0x1.c7ba88988c933p+0, 0x1.c8f6d9406e7b5p+0, 0x1.ca3405751c4dbp+0,
    0x1.cb720dcef9069p+0, 0x1.ccb0f2e6d1675p+0, 0x1.cdf0b555dc3fap+0, 0x1.cf3155b5bab74p+0,
    0x1.d072d4a07897cp+0, 0x1.d1b532b08c968p+0, 0x1.d2f87080d89f2p+0, 0x1.d43c8eacaa1d6p+0,
    0x1.d5818dcfba487p+0, 0x1.d6c76e862e6d3p+0, 0x1.d80e316c98398p+0, 0x1.d955d71ff6075p+0,
    0x1.da9e603db3285p+0, 0x1.dbe7cd63a8315p+0, 0x1.dd321f301b460p+0, 0x1.de7d5641c0658p+0,
    0x1.dfc97337b9b5fp+0, 0x1.e11676b197d17p+0, 0x1.e264614f5a129p+0, 0x1.e3b333b16ee12p+0,
    0x1.e502ee78b3ff6p+0, 0x1.e653924676d76p+0, 0x1.e7a51fbc74c83p+0, 0x1.e8f7977cdb740p+0,
    0x1.ea4afa2a490dap+0, 0x1.eb9f4867cca6ep+0, 0x1.ecf482d8e67f1p+0, 0x1.ee4aaa2188510p+0,
    0x1.efa1bee615a27p+0, 0x1.f0f9c1cb6412ap+0, 0x1.f252b376bba97p+0, 0x1.f3ac948dd7274p+0,
    0x1.f50765b6e4540p+0, 0x1.f6632798844f8p+0, 0x1.f7bfdad9cbe14p+0, 0x1.f91d802243c89p+0,
    0x1.fa7c1819e90d8p+0, 0x1.fbdba3692d514p+0, 0x1.fd3c22b8f71f1p+0, 0x1.fe9d96b2a23d9p+0};
constexpr double kExpInvStep = 0x1.71547652b82fep+8;   // 256/ln2
constexpr double kExpStep = 0x1.62e42fefa39efp-9;      // ln2/256

// every thread of the CTA calls this once, before a __syncthreads(), to stage the table
__device__ __forceinline__ void exp_tab_stage(double* tab) {
  for (int i = threadIdx.x; i < kExpTabN; i += blockDim.x) tab[i] = kExpTab[i];
}

__device__ __forceinline__ double exp_tb(double x, const double* __restrict__ tab) {
  const double kMagic = 6755399441055744.0;  // 1.5 * 2^52
  double t = fma(x, kExpInvStep, kMagic);
  const int n = __double2loint(t);
  t -= kMagic;
  const double r = fma(-t, kExpStep, x);
  const double T = tab[n & (kExpTabN - 1)];
  double p = fma(r, 4.1666666666666664e-02, 1.6666666666666666e-01);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(T, p * r, T);
  return __hiloint2double(__double2hiint(p) + ((n >> 8) << 20), __double2loint(p));
}

// N independent arguments, stage by stage (see the note on ordered FP64 ops above)
template <int N>
__device__ __forceinline__ void exp_tb_vec(const double (&x)[N], double (&y)[N],
                                           const double* __restrict__ tab) {
  const double kMagic = 6755399441055744.0;
  double t[N], r[N], p[N], T[N];
  int n[N];
#pragma unroll
  for (int a = 0; a < N; ++a) MWRT_FMA(t[a], x[a], kExpInvStep, kMagic);
#pragma unroll
  for (int a = 0; a < N; ++a) {
    n[a] = __double2loint(t[a]);
    T[a] = tab[n[a] & (kExpTabN - 1)];
    MWRT_ADD(t[a], t[a], -kMagic);
  }
#pragma unroll
  for (int a = 0; a < N; ++a) MWRT_FMA(r[a], t[a], -kExpStep, x[a]);
#pragma unroll
  for (int a = 0; a < N; ++a) MWRT_FMA(p[a], r[a], 4.1666666666666664e-02, 1.6666666666666666e-01);
#pragma unroll
  for (int a = 0; a < N; ++a) MWRT_FMA(p[a], p[a], r[a], 0.5);
#pragma unroll
  for (int a = 0; a < N; ++a) MWRT_FMA(p[a], p[a], r[a], 1.0);
#pragma unroll
  for (int a = 0; a < N; ++a) MWRT_MUL(p[a], p[a], r[a]);
#pragma unroll
  for (int a = 0; a < N; ++a) MWRT_FMA(p[a], T[a], p[a], T[a]);
#pragma unroll
  for (int a = 0; a < N; ++a)
    y[a] = __hiloint2double(__double2hiint(p[a]) + ((n[a] >> 8) << 20), __double2loint(p[a]));
}

// ---- table-driven natural log: 14 FP64-pipe instructions ----------------------------------------
// x = 2^e m, m in [1,2); i = top 6 mantissa bits.  Intervals i >= 27 (m >= 1.42) are taken as m/2 with
// e+1, so the argument of the table step lies in [0.71, 1.42) and log x never cancels against e ln2.
// Table (c_i, -log c_i): c_i ~ 1/m_centre to 12 bits, so r = fma(m, c_i, -1) is exact-rounded and
// |r| <= 1/128; the two intervals touching 1 use c = 1 (r = m - 1 exactly, table term 0), so the
// result keeps full RELATIVE accuracy as x -> 1.  log1p(r) to degree 9 (|r| <= 1/64: r^9/10 < 6e-18).
// Domain: normal positive x (callers substitute a stand-in and select afterwards).
static __device__ const double kLogTab64[128] = {
    0x1.0000000000000p+0, 0x0.0p+0, 0x1.f440000000000p-1, 0x1.7c61b1cf5dee0p-6,
    0x1.ecc0000000000p-1, 0x1.39f07ba0ebd62p-5, 0x1.e580000000000p-1, 0x1.b35dd9b58baadp-5,
    0x1.de60000000000p-1, 0x1.163d6ef957a03p-4, 0x1.d780000000000p-1, 0x1.518874226130ap-4,
    0x1.d0c0000000000p-1, 0x1.8c985e9b9ec84p-4, 0x1.ca40000000000p-1, 0x1.c6494a2e418a6p-4,
    0x1.c400000000000p-1, 0x1.fe89139dbd566p-4, 0x1.bde0000000000p-1, 0x1.1b35ae3b81dbfp-3,
    0x1.b7e0000000000p-1, 0x1.36f4c27577593p-3, 0x1.b200000000000p-1, 0x1.527e5e4a1b58dp-3,
    0x1.ac60000000000p-1, 0x1.6d35fee52b83bp-3, 0x1.a6e0000000000p-1, 0x1.87ad07c493478p-3,
    0x1.a160000000000p-1, 0x1.a27cc30640ecbp-3, 0x1.9c20000000000p-1, 0x1.bc69684aee63ep-3,
    0x1.9700000000000p-1, 0x1.d60a17f903515p-3, 0x1.9200000000000p-1, 0x1.ef5ade4dcffe6p-3,
    0x1.8d40000000000p-1, 0x1.03d95a1d67686p-2, 0x1.8860000000000p-1, 0x1.107e404ab0f81p-2,
    0x1.83c0000000000p-1, 0x1.1ca28c64bae54p-2, 0x1.7f40000000000p-1, 0x1.2896a13e086a4p-2,
    0x1.7ae0000000000p-1, 0x1.34585a594b8adp-2, 0x1.7680000000000p-1, 0x1.403d086cea79cp-2,
    0x1.7240000000000p-1, 0x1.4becf95d97913p-2, 0x1.6e20000000000p-1, 0x1.5765f1749da6bp-2,
    0x1.6a20000000000p-1, 0x1.62a5afc06121fp-2, 0x1.6620000000000p+0, -0x1.57c2f53b05209p-2,
    0x1.6240000000000p+0, -0x1.4c9f09e152c3cp-2, 0x1.5e80000000000p+0, -0x1.41b941cce0beep-2,
    0x1.5ac0000000000p+0, -0x1.36b5776bc1117p-2, 0x1.5720000000000p+0, -0x1.2bf287cc41362p-2,
    0x1.53a0000000000p+0, -0x1.2172d5c2abbbbp-2, 0x1.5020000000000p+0, -0x1.16d74c2e8ac66p-2,
    0x1.4ca0000000000p+0, -0x1.0c1f55d88a8a1p-2, 0x1.4960000000000p+0, -0x1.021164a9290ddp-2,
    0x1.4600000000000p+0, -0x1.ef0adcbdc5936p-3, 0x1.42e0000000000p+0, -0x1.db50da24bd322p-3,
    0x1.3fc0000000000p+0, -0x1.c765b9e4d68d6p-3, 0x1.3ca0000000000p+0, -0x1.b34885022e81ep-3,
    0x1.39a0000000000p+0, -0x1.9fc93e5409c31p-3, 0x1.36a0000000000p+0, -0x1.8c19fe2982058p-3,
    0x1.33a0000000000p+0, -0x1.7839d62023552p-3, 0x1.30e0000000000p+0, -0x1.65d5e99cc3c47p-3,
    0x1.2e00000000000p+0, -0x1.526e5e3a1b438p-3, 0x1.2b40000000000p+0, -0x1.3fb25a59528cbp-3,
    0x1.2880000000000p+0, -0x1.2cca0f5f5f251p-3, 0x1.25e0000000000p+0, -0x1.1a93b7d42f611p-3,
    0x1.2340000000000p+0, -0x1.08338affa282ap-3, 0x1.20c0000000000p+0, -0x1.ed1794e83780cp-4,
    0x1.1e20000000000p+0, -0x1.c7af7a8472d76p-4, 0x1.1bc0000000000p+0, -0x1.a58b60c2b2324p-4,
    0x1.1940000000000p+0, -0x1.814be23f8c036p-4, 0x1.16e0000000000p+0, -0x1.5e8fa4d8591c7p-4,
    0x1.1480000000000p+0, -0x1.3b87598b1b6eep-4, 0x1.1240000000000p+0, -0x1.1a0fba1bf8a52p-4,
    0x1.0fe0000000000p+0, -0x1.ecdf0f87b6168p-5, 0x1.0dc0000000000p+0, -0x1.ac97221711f8dp-5,
    0x1.0b80000000000p+0, -0x1.67f94f094bd98p-5, 0x1.0960000000000p+0, -0x1.26a32a86d2f72p-5,
    0x1.0740000000000p+0, -0x1.c98d18d00c814p-6, 0x1.0520000000000p+0, -0x1.44c28d451662cp-6,
    0x1.0300000000000p+0, -0x1.7dc475f810a77p-7, 0x1.0000000000000p+0, 0x0.0p+0};

__device__ __forceinline__ void log_tab_stage(double* tab) {
  for (int i = threadIdx.x; i < 128; i += blockDim.x) tab[i] = kLogTab64[i];
}

__device__ __forceinline__ double log_tb(double x, const double* __restrict__ tab) {
  int hi = __double2hiint(x);
  const int lo = __double2loint(x);
  int e = (hi >> 20) - 1023;
  const int i = (hi >> 14) & 63;
  const bool big = i >= 27;
  hi = (hi & 0x000fffff) | (big ? 0x3fe00000 : 0x3ff00000);   // m or m/2
  e = big ? e + 1 : e;
  const double m = __hiloint2double(hi, lo);
  const double2 cl = *reinterpret_cast<const double2*>(tab + 2 * i);
  const double r = fma(m, cl.x, -1.0);
  const double r2 = r * r;
  double q = 1.0 / 9.0;
  q = fma(q, r, -1.0 / 8.0);
  q = fma(q, r, 1.0 / 7.0);
  q = fma(q, r, -1.0 / 6.0);
  q = fma(q, r, 1.0 / 5.0);
  q = fma(q, r, -1.0 / 4.0);
  q = fma(q, r, 1.0 / 3.0);
  q = fma(q, r, -0.5);
  const double ed = (double)e;
  const double tail = fma(r2, q, ed * 1.90821492927058770002e-10);
  const double head = fma(ed, 6.93147180369123816490e-01, cl.y);
  return head + (r + tail);
}

// N reciprocals with the single quadratic Newton step of rcp_nr1, staged
template <int N>
__device__ __forceinline__ void rcp_nr1_vec(const double (&x)[N], double (&y)[N]) {
  double r[N], e[N];
#pragma unroll
  for (int a = 0; a < N; ++a) asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r[a]) : "d"(x[a]));
#pragma unroll
  for (int a = 0; a < N; ++a) MWRT_FMA(e[a], -x[a], r[a], 1.0);
#pragma unroll
  for (int a = 0; a < N; ++a) MWRT_FMA(y[a], r[a], e[a], r[a]);
}

// ---- square root without the IEEE slow path -------------------------------------------------------
// MUFU.RSQ64H seed + two coupled Newton steps + one residual correction (<= 1 ulp) for normal x > 0.
__device__ __forceinline__ double sqrt_f(double x) {  // x > 0, normal
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double h = 0.5 * y, s = x * y;
  double e = fma(-s, h, 0.5);
  s = fma(s, e, s);
  h = fma(h, e, h);
  e = fma(-s, h, 0.5);
  s = fma(s, e, s);
  h = fma(h, e, h);
  return fma(fma(-s, s, x), h, s);
}

// ---- branch-free atan2 ---------------------------------------------------------------------------
// Principal value of atan2(y, x) for finite (x, y) != (0, 0), no library call, no branch: the ratio
// t = min(|x|,|y|)/max(|x|,|y|) in [0,1] is reduced around c in {0, 1/2, 1} with
// atan t = atan c + atan((t-c)/(1+tc)), |u| <= 1/4, and atan u is its odd series to u^27
// (truncation u^28/29 < 8e-18 relative); quadrants by selects.  ~40 FP64-pipe instructions, <= 2 ulp.
__device__ __forceinline__ double atan2_nb(double y, double x) {
  const double ax = fabs(x), ay = fabs(y);
  const double mx = fmax(ax, ay), mn = fmin(ax, ay);
  const double t = div_fast(mn, mx);
  const bool b1 = t > 0.25, b2 = t > 0.75;
  const double c = b2 ? 1.0 : (b1 ? 0.5 : 0.0);
  const double ac = b2 ? 0.78539816339744830962 : (b1 ? 0.46364760900080611621 : 0.0);
  const double u = div_fast(t - c, fma(t, c, 1.0));
  const double u2 = u * u;
  double q = -1.0 / 27.0;
  q = fma(q, u2, 1.0 / 25.0);
  q = fma(q, u2, -1.0 / 23.0);
  q = fma(q, u2, 1.0 / 21.0);
  q = fma(q, u2, -1.0 / 19.0);
  q = fma(q, u2, 1.0 / 17.0);
  q = fma(q, u2, -1.0 / 15.0);
  q = fma(q, u2, 1.0 / 13.0);
  q = fma(q, u2, -1.0 / 11.0);
  q = fma(q, u2, 1.0 / 9.0);
  q = fma(q, u2, -1.0 / 7.0);
  q = fma(q, u2, 1.0 / 5.0);
  q = fma(q, u2, -1.0 / 3.0);
  double r = ac + fma(u * u2, q, u);
  r = (ay > ax) ? 1.57079632679489661923 - r : r;
  r = (x < 0.0) ? 3.14159265358979323846 - r : r;
  return (y < 0.0) ? -r : r;
}

// ---- minimal complex arithmetic (speed-dependent line shape, liquid permittivity) -----------
struct cplx {
  double re, im;
};
__device__ __forceinline__ cplx cmake(double a, double b) { return cplx{a, b}; }
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return cplx{a.re + b.re, a.im + b.im}; }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return cplx{a.re - b.re, a.im - b.im}; }
__device__ __forceinline__ cplx cmul(cplx a, cplx b) {
  return cplx{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
}
__device__ __forceinline__ cplx cscale(cplx a, double s) { return cplx{a.re * s, a.im * s}; }
__device__ __forceinline__ cplx cconj(cplx a) { return cplx{a.re, -a.im}; }
__device__ __forceinline__ cplx cdiv(cplx a, cplx b) {
  // Smith's algorithm (what C99/NumPy use) to stay close to the oracle's rounding
  if (fabs(b.re) >= fabs(b.im)) {
    double r = b.im / b.re;
    double d = 1.0 / (b.re + b.im * r);
    return cplx{(a.re + a.im * r) * d, (a.im - a.re * r) * d};
  } else {
    double r = b.re / b.im;
    double d = 1.0 / (b.re * r + b.im);
    return cplx{(a.re * r + a.im) * d, (a.im * r - a.re) * d};
  }
}

// Goff-Gratch saturation vapour pressure over water [mb] (RTEquation.vapor), with the
// table-driven exp/log (10^x = exp(x ln 10), log10 y = ln y / ln 10) instead of libm's exp10/log10 and
// their slow paths.  tk must be a physical temperature (callers substitute a stand-in otherwise).
__device__ __forceinline__ double goff_gratch_es(double tk, const double* __restrict__ etab,
                                                 const double* __restrict__ ltab) {
  constexpr double kLn10 = 2.302585092994045684, kInvLn10 = 0.434294481903251828;
  const double y = div_fast(373.16, tk);
  const double es = -7.90298 * (y - 1.0) + 5.02808 * (log_tb(y, ltab) * kInvLn10) -
                    1.3816e-07 * (exp_tb((11.344 * (1.0 - div_fast(1.0, y))) * kLn10, etab) - 1.0) +
                    0.0081328 * (exp_tb((-3.49149 * (y - 1.0)) * kLn10, etab) - 1.0) +
                    3.0057148979490314;  // log10(1013.246)
  return exp_tb(es * kLn10, etab);
}

}  // namespace mwrt

#define MWRT_CUDA_OK(call)                         \
  do {                                             \
    cudaError_t e__ = (call);                      \
    if (e__ != cudaSuccess) return mwrt::fail_cuda(e__, #call); \
  } while (0)

namespace mwrt {
int fail_cuda(cudaError_t e, const char* what);
int fail_arg(int code, const char* what);
}  // namespace mwrt
