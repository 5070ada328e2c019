// mwrt_rte_kernel.cuh -- kernel 3 (k_rte) and its launcher family, included by mwrt_rte.cu (plain
// result stores) and mwrt_rte_fanout.cu (stores fanned out to other GPUs) so the two template families
// compile side by side.
//
// Kernel 3 (k_rte): RTEquation.exponential_integration (layer optical depths), RTEquation.planck
//   (downwelling recursion + cosmic background, or upwelling), RTEquation.bright, the cloud
//   liquid/ice absorption of the cloudy levels (RTEquation.cloudy_absorption, evaluated inline) and the
//   channel band-pass average.  SURVEY.md A.7/A.8/A.10; reached by the reference through `rte.execute()`
//   at /root/reference/python_src/proc/PyRTlib_processing.py:126.
//
// Mapping: thread = (profile, frequency); it walks the layers once, in the oracle's order, and carries
// NA viewing angles in registers -- the layer-mean absorption (2-4 log) and the Planck term (1 exp)
// are computed once per layer and shared by all angles, and the NA exp(-tau) chains give the FP64 pipe
// NA-way instruction-level parallelism.  One exp per (layer, angle): the transmission to the observer
// is carried as a running product instead of PyRTlib's second exp(-tauprof).
#pragma once
#include "mwrt_common.cuh"
#include "mwrt_cloud.cuh"
#include "mwrt_rte_args.h"
#include <type_traits>

namespace mwrt {

// One layer of exponential_integration without the ds factor: x0 = lower level,
// x1 = upper level.  Branch-free (selects) so the caller's layer loop stays one basic block:
//   negative input -> NaN and `neg` (PyRTlib raises ValueError; the profile gets MWRT_ST_NEG_ABS)
//   |x1-x0| < 1e-9 -> x1;  a zero endpoint -> mean (zeroflg) or 0;  else (x1-x0)/log(x1/x0).
__device__ __forceinline__ double layer_mean(double x0, double x1, bool zeroflg, bool& neg,
                                             const double* __restrict__ ltab) {
  const bool negv = (x0 < 0.0) || (x1 < 0.0);
  const bool close = fabs(x1 - x0) < 1e-09;
  const bool zero = (x0 == 0.0) || (x1 == 0.0);
  const bool use_log = !(negv || close || zero) && (x0 > 0.0) && (x1 > 0.0);  // false for NaN too
  const double a = use_log ? x1 : 2.0, bq = use_log ? x0 : 1.0;              // keeps log_tb's domain
  // reciprocal-multiply (<= 1.5 ulp each) instead of corrected divisions: 4 FP64 instructions fewer
  const double lm = (x1 - x0) * rcp_fast(log_tb(a * rcp_fast(bq), ltab));
  double r = zero ? (zeroflg ? (x1 + x0) * 0.5 : 0.0) : lm;
  r = close ? x1 : r;
  r = (negv || !(use_log || close || zero)) ? nan("") : r;
  neg = neg || negv;
  return r;
}

// modified Planck function 1/(exp(hvk/T) - 1); hvk/T is O(1e-2) for microwave frequencies.
// T is a physical temperature here (kernel 1 flags anything outside 50..1000 K), hvk/T < 700.
__device__ __forceinline__ double planck_mod(double hvk, double t, const double* __restrict__ etab) {
  const double x = hvk * rcp_fast(t);
  return rcp_fast(exp_tb(fmax(fmin(x, 700.0), -700.0), etab) - 1.0);
}
// bright() without library calls or branches: b <= 0 -> 0, NaN -> NaN, else hvk / log(1 + 1/b)
__device__ __forceinline__ double bright_nb(double hvk, double b, const double* __restrict__ ltab) {
  const bool pos = b > 0.0;
  const double x = 1.0 + rcp_fast(pos ? b : 1.0);
  const double t = hvk * rcp_fast(log_tb(x, ltab));
  return pos ? t : (isnan(b) ? b : 0.0);
}

// layer_mean for two strictly positive endpoints (the common case): same operations, no selects
// beyond the |x1-x0| < 1e-9 rule
__device__ __forceinline__ double layer_mean_pos(double x0, double x1, const double* __restrict__ ltab) {
  const double dx = x1 - x0;
  const double lm = dx * rcp_fast(log_tb(x1 * rcp_fast(x0), ltab));
  return (fabs(dx) < 1e-09) ? x1 : lm;
}

// PLANE = plane-parallel path: ds(a, layer) = dz(layer) * amass(a) is formed in registers from z and
// a per-angle air-mass table (no ds array, no slant-path kernel); otherwise ds is read from (P,A,L).
// FROM_SAT (upwelling, levels walked top-down) is a template parameter so the lower/upper level of a
// layer is chosen at compile time instead of by per-layer selects.
// measured on B200 (profiles/r01b_*): the plane-parallel variant is fastest capped at 168 registers
// (3 CTAs/SM).
#ifndef MWRT_RTE_MINB
#define MWRT_RTE_MINB 3
#endif
#ifndef MWRT_RTE_MINB_CLOUDY
#define MWRT_RTE_MINB_CLOUDY 3   // measured: 168 registers without spills, config 3 share -1 %
#endif
#ifndef MWRT_RTE_MINB_NA1
#define MWRT_RTE_MINB_NA1 MWRT_RTE_MINB
#endif
#ifndef MWRT_RTE_MINB_GENERAL
#define MWRT_RTE_MINB_GENERAL 2  // ray-traced path: per-angle wet and total sums live in registers
                                 // (measured: 3 CTAs/SM at 168 registers spills and is 6 % slower)
#endif
// FANOUT selects the result-store form at compile time (plain stores, or the warp-staged stores that
// fan out to other GPUs): the staged epilogue costs the plain kernel nothing, not even uniform registers.
// BP: the channel band-pass average (BASELINE.json config 3) is reduced INSIDE the kernel.  A CTA then
// covers `cpb` whole channels of one profile (thread = one sub-frequency, <= kRteBpMaxThreads threads);
// after the layer walk each field goes through shared memory once and thread (channel, angle) adds its
// S sub-frequencies in index order with one fma each -- the same chain the separate k_bandpass runs, so
// the result is bit-identical and deterministic, and no monochromatic (P,F,A) field ever reaches HBM.
template <int NA, bool CLOUDY, bool PLANE, bool FROM_SAT, bool FANOUT, bool BP>
__global__ void __launch_bounds__(BP ? kRteBpMaxThreads : kRteThreads,
                                  BP ? 1 : (!PLANE ? MWRT_RTE_MINB_GENERAL
                                                   : (CLOUDY ? MWRT_RTE_MINB_CLOUDY
                                                             : (NA <= 2 ? MWRT_RTE_MINB_NA1 : MWRT_RTE_MINB))))
k_rte(const RteArgs g, const Fanout fo) {
  const int P = g.P, L = g.L, F = g.F, A = g.A;
  const double* __restrict__ tk = g.tk;
  const double* __restrict__ awet = g.awet;
  const double* __restrict__ adry = g.adry;
  const double* __restrict__ aliq = g.aliq;
  const double* __restrict__ aice = g.aice;
  const double* __restrict__ ds = g.ds;
  const double* __restrict__ z = g.z;
  const double* __restrict__ amass_tab = g.amass;
  const double* __restrict__ frq = g.frq;
  const double emis = g.emis;
  double* __restrict__ tbtotal = g.tbtotal;
  double* __restrict__ tbatm = g.tbatm;
  double* __restrict__ tmr = g.tmr;
  double* __restrict__ tauwet = g.tauwet;
  double* __restrict__ taudry = g.taudry;
  double* __restrict__ tauliq = g.tauliq;
  double* __restrict__ tauice = g.tauice;
  double* __restrict__ tau_layer = g.tau_layer;
  const int tl_angle = g.tl_angle;
  int* __restrict__ status = g.status;

  __shared__ double etab[kExpTabN];        // exp_tb / log_tb tables (mwrt_common.cuh)
  __shared__ __align__(16) double ltab[128];
  exp_tab_stage(etab);
  log_tab_stage(ltab);
  __syncthreads();
  // threads past the end of the (profile, frequency) range redo the last valid pair and store
  // nothing: every lane stays alive for the cooperative result stores at the end
  const long long npf = (long long)P * F;
  long long idx_raw, p;
  int jf, bp_c = 0;
  bool active;
  if constexpr (BP) {
    const int ngrp = (g.bpC + g.bpCpb - 1) / g.bpCpb;
    p = blockIdx.x / ngrp;
    const int cl = threadIdx.x / g.bpS;
    bp_c = (int)(blockIdx.x - p * ngrp) * g.bpCpb + cl;
    active = (cl < g.bpCpb) && (bp_c < g.bpC);
    jf = active ? bp_c * g.bpS + (threadIdx.x - cl * g.bpS) : F - 1;
    idx_raw = p * F + jf;
  } else {
    idx_raw = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    active = idx_raw < npf;
    const long long idc = active ? idx_raw : npf - 1;
    p = idc / F;
    jf = (int)(idc - p * F);
  }
  const long long idx = p * F + jf;
  const int a0 = blockIdx.y * NA;
  const long long row = idx * (long long)L;  // (p*F + jf)*L
  const double* tkp = tk + p * L;
  const double* zp = PLANE ? z + p * L : nullptr;
  const double* dsp = PLANE ? nullptr : ds + (p * A + a0) * (long long)L;
  const double fq = frq[jf];
  const double hvk = (fq * 1e9) * kPlanckH / kBoltzmannK;
  // cloud absorption evaluated inline (no (P,F,L) liquid/ice arrays): lanes of a warp are frequencies
  // of ONE profile, so "this level holds liquid" is warp-uniform
  const bool cld_inline = CLOUDY && (aliq == nullptr);
  const double* __restrict__ dlq = cld_inline ? g.denliq + p * L : nullptr;
  const double* __restrict__ dic = cld_inline ? g.denice + p * L : nullptr;
  const double ice_c = CLOUDY ? ice_coef(fq) : 0.0;

  double batm[NA], trans[NA], amass[NA];  // (amass in shared memory measured slower)
  // species sums: per angle in the general case, one zenith-path sum (scaled at the end) if PLANE
  constexpr int NS = PLANE ? 1 : NA;
  double sw[NS], sd[NS], sl[NS], si[NS], tauprof[NS];
#pragma unroll
  for (int a = 0; a < NA; ++a) {
    batm[a] = 0.0;
    trans[a] = 1.0;
    amass[a] = (PLANE && a0 + a < A) ? amass_tab[a0 + a] : 0.0;
  }
#pragma unroll
  for (int a = 0; a < NS; ++a) { sw[a] = 0.0; sd[a] = 0.0; sl[a] = 0.0; si[a] = 0.0; tauprof[a] = 0.0; }

  // PLANE: one clamp per layer on the zenith optical depth instead of one per angle.  exp_tb needs
  // tau < 700; capping the zenith value at 700/max(amass) changes exp(-tau) only where it is < 1e-21
  // relative to the terms it is added to.
  double tcap = 700.0;
  if (PLANE) {
    double amax = 1.0;
#pragma unroll
    for (int a = 0; a < NA; ++a) amax = fmax(amax, fabs(amass[a]));
    tcap = 700.0 / amax;
  }
  double tl_amass = 0.0;   // PLANE: air mass of the angle whose per-layer optical depths are wanted
#pragma unroll
  for (int a = 0; a < NA; ++a)
    if (a0 + a == tl_angle) tl_amass = amass[a];

  bool neg = false;
  double tsum = 0.0;       // NaN detector for the temperatures (NaN in absorption or geometry shows
                           // up in tauprof): checked once after the loop instead of per layer
  const bool want_tl = (tau_layer != nullptr) && (tl_angle >= a0) && (tl_angle < a0 + NA);
  if (want_tl && active) tau_layer[row] = 0.0;
  constexpr int dir = FROM_SAT ? -1 : 1;
  const int lev0 = FROM_SAT ? L - 1 : 0;   // level sequence starts at the observer's end
  const double zref = PLANE ? zp[0] : 0.0;

  // Software pipeline over the level sequence lev0, lev0+dir, ...:
  //   A = level reached last, B = level being reached (already in registers), C = prefetched.
  // The scalars of the NEXT layer (2-4 layer means, 1 Planck term: serial log/exp chains) are
  // computed in the same basic block as the CURRENT layer's NA-angle stage, so they overlap.
  struct Lvl { double w, d, l, i, t, z; };
  struct Lay { double xw, xd, xl, xi, xtot, bfar, dz; };
  auto load = [&](int lv) {
    Lvl v;
    v.w = awet[row + lv];
    v.d = adry[row + lv];
    v.t = tkp[lv];
    v.l = 0.0;
    v.i = 0.0;
    if (CLOUDY) {
      if (cld_inline) {
        const double dl = dlq[lv], di = dic[lv];
        if (dl > 0.0 || isnan(dl)) v.l = liq_abs_unit(g.liq_model, v.t, fq, etab, ltab) * dl;
        if (di > 0.0 || isnan(di)) v.i = ice_abs_c(ice_c, di);
      } else {
        v.l = aliq[row + lv];
        v.i = aice[row + lv];
      }
    }
    v.z = PLANE ? zp[lv] - zref : 0.0;
    return v;
  };
  // GENERIC = false: wet and dry absorption strictly positive at both levels (decided by the caller)
  auto prep = [&](auto generic, const Lvl& nr, const Lvl& fr, bool& ng) {
    constexpr bool GENERIC = decltype(generic)::value;
    // nr = level nearer to the observer, fr = farther; lower/upper depends on the direction
    const Lvl& lo = FROM_SAT ? fr : nr;
    const Lvl& up = FROM_SAT ? nr : fr;
    Lay y;
    if (GENERIC) {
      y.xw = layer_mean(lo.w, up.w, true, ng, ltab);
      y.xd = layer_mean(lo.d, up.d, true, ng, ltab);
    } else {
      y.xw = layer_mean_pos(lo.w, up.w, ltab);
      y.xd = layer_mean_pos(lo.d, up.d, ltab);
    }
    y.xtot = y.xw + y.xd;
    y.xl = 0.0;
    y.xi = 0.0;
    if (CLOUDY) {
      // cloud-free level pairs (the large majority) skip both means: exponential_integration gives
      // exactly 0 for two zero endpoints
      if (lo.l != 0.0 || up.l != 0.0 || lo.i != 0.0 || up.i != 0.0) {
        y.xl = layer_mean(lo.l, up.l, false, ng, ltab);
        y.xi = layer_mean(lo.i, up.i, false, ng, ltab);
      }
      y.xtot = y.xtot + y.xi + y.xl;
    }
    y.bfar = planck_mod(hvk, fr.t, etab);
    y.dz = up.z - lo.z;
    return y;
  };
  auto all_pos = [](const Lvl& a, const Lvl& b) {
    return (a.w > 0.0) && (b.w > 0.0) && (a.d > 0.0) && (b.d > 0.0);
  };

  Lvl lvA = load(lev0);
  tsum = lvA.t;
  double bnear = planck_mod(hvk, lvA.t, etab);
  Lvl lvB = load(lev0 + dir);
  Lay cur = prep(std::true_type{}, lvA, lvB, neg);
  lvA = lvB;
  lvB = load(L > 2 ? lev0 + 2 * dir : lev0 + dir);

  // one step: the NA-angle stage of the current layer + the scalars of the next one
  auto step_body = [&](auto generic, const int step) {
    const int clev = lev0 + step * dir;          // level reached in this step
    const int j = FROM_SAT ? clev + 1 : clev;    // layer index (upper level of the layer)
    // next layer's scalars (discarded after the last step) + prefetch of the level after it
    bool ng2 = false;
    const Lay nxt = prep(generic, lvA, lvB, ng2);
    const bool more = step + 1 < L;
    neg = neg || (more && ng2);
    tsum += lvA.t;
    lvA = lvB;
    lvB = load(step + 2 < L ? lev0 + (step + 2) * dir : clev);

    const double xw = cur.xw, xd = cur.xd, xl = cur.xl, xi = cur.xi, xtot = cur.xtot;
    const double bfar = cur.bfar, dz = cur.dz;
    if (PLANE) {
      sw[0] = fma(xw, dz, sw[0]);
      sd[0] = fma(xd, dz, sd[0]);
      if (CLOUDY) { sl[0] = fma(xl, dz, sl[0]); si[0] = fma(xi, dz, si[0]); }
      tauprof[0] = fma(xtot, dz, tauprof[0]);
    }
    // per-angle work, written stage by stage over the NA angles (see exp_tb_vec)
    double tau[NA], ntau[NA], ex[NA], den[NA], rden[NA];
    if (PLANE) {
      const double t0 = xtot * dz;                 // zenith optical depth of the layer
      const double t0n = -fmin(t0, tcap);          // NaN is caught through tauprof after the loop
#pragma unroll
      for (int a = 0; a < NA; ++a) {
        ntau[a] = t0n * amass[a];
        tau[a] = 0.0;
      }
      if (want_tl && active) tau_layer[row + j] = t0 * tl_amass;
    } else {
      // per-angle sums: wet (+ liquid, ice) and the total; the dry sum is total - the others at the end
      // (one subtraction per output instead of NA fused multiply-adds per layer and NA live registers)
      // (prefetching the next layer's path lengths one step ahead was measured 7 % SLOWER: 20 more live
      // registers in a kernel that already sits at 244)
#pragma unroll
      for (int a = 0; a < NA; ++a) {
        const double dsa = (a0 + a < A) ? dsp[(long long)a * L + j] : 0.0;
        sw[a] = fma(xw, dsa, sw[a]);
        if (CLOUDY) { sl[a] = fma(xl, dsa, sl[a]); si[a] = fma(xi, dsa, si[a]); }
        tau[a] = xtot * dsa;
        tauprof[a] += tau[a];
      }
#pragma unroll
      for (int a = 0; a < NA; ++a) ntau[a] = fmax(-tau[a], -700.0);
    }
    exp_tb_vec<NA>(ntau, ex, etab);
#pragma unroll
    for (int a = 0; a < NA; ++a) den[a] = 1.0 + ex[a];
    rcp_nr1_vec<NA>(den, rden);
#pragma unroll
    for (int a = 0; a < NA; ++a) {
      // B(layer) * (trans_before - trans_after): the weight (1 - e^-tau) e^-tau_above of RTEquation.planck
      const double boftlay = fma(bfar, ex[a], bnear) * rden[a];
      const double tn = trans[a] * ex[a];
      batm[a] = fma(boftlay, trans[a] - tn, batm[a]);
      trans[a] = tn;
    }
    if (want_tl && !PLANE && active) {
      double tl = 0.0;
#pragma unroll
      for (int a = 0; a < NA; ++a)
        if (a0 + a == tl_angle) tl = tau[a];
      tau_layer[row + j] = tl;
    }
    bnear = bfar;
    cur = nxt;
  };

  // the two copies of the body differ only in how the next layer's means are formed; keeping each a
  // single basic block lets the serial log/exp chains overlap the NA-angle stage.  Unrolling by two
  // removes the register rotation of the pipeline (cur <- nxt, A <- B) in the plane-parallel kernel;
  // the general kernel is larger and measured slower unrolled (instruction cache).
  if (PLANE) {
#pragma unroll 2
    for (int step = 1; step < L; ++step) {
      if (all_pos(lvA, lvB)) step_body(std::false_type{}, step);
      else step_body(std::true_type{}, step);
    }
  } else {
#pragma unroll 1
    for (int step = 1; step < L; ++step) {
      if (all_pos(lvA, lvB)) step_body(std::false_type{}, step);
      else step_body(std::true_type{}, step);
    }
  }

  bool anynan = isnan(tsum);
#pragma unroll
  for (int a = 0; a < NS; ++a) anynan = anynan || isnan(tauprof[a]);
  const bool flagged = status && (status[p] & (MWRT_ST_NAN_INPUT | MWRT_ST_RANGE));  // earlier kernels
  const bool bad = neg || anynan || flagged;
  if (status && neg && active) atomicOr(&status[p], MWRT_ST_NEG_ABS);
  if (status && anynan && !neg && !flagged && jf == 0 && a0 == 0 && active)
    atomicOr(&status[p], MWRT_ST_NAN_INPUT);
  const double bbg = FROM_SAT ? emis * planck_mod(hvk, tkp[0], etab) : planck_mod(hvk, kTCosmic, etab);

  // cosmic background / T_mr (RTEquation.planck) and the inverse Planck function (bright), with the
  // branch-free exp/log/reciprocal helpers: NA angles x 3 outputs would otherwise inline ~6000
  // instructions of library slow paths here
  // (evaluated inside the store helpers below, per field and angle: nothing is kept in arrays)
  auto et_of = [&](int a) {
    const double tp = PLANE ? tauprof[0] * amass[a] : tauprof[a];
    return (tp < kExpMax) ? exp_tb(fmax(-tp, -700.0), etab) : -1.0;     // -1: opaque path
  };
  const double qn = nan("");
  auto v_tot = [&](int a) {
    const double et = et_of(a);
    return bad ? qn : bright_nb(hvk, (et >= 0.0) ? fma(bbg, et, batm[a]) : batm[a], ltab);
  };
  auto v_atm = [&](int a) { return bad ? qn : bright_nb(hvk, batm[a], ltab); };
  auto v_mr = [&](int a) {
    const double et = et_of(a);
    return bad ? qn : bright_nb(hvk, (et >= 0.0) ? batm[a] * rcp_fast(1.0 - et) : batm[a], ltab);
  };

  // Result stores.
  //  plain   a thread owns A consecutive doubles per field and stores each value directly (the L2
  //          merges them well enough);
  //  FANOUT  a warp owns 32*A consecutive doubles: when this kernel covers all angles (A == NA) and the
  //          field is 16-byte aligned, the warp transposes them through its private slice of shared
  //          memory (per HALF warp: 5 KB per CTA) and writes 16 bytes per lane, fully coalesced --
  //          which is what makes the NVLink multicast form (Fanout mode 1: one `multimem.st.v2` per 16
  //          bytes, replicated to every GPU by NVSwitch) as cheap as a local store;
  //  BP      every field goes through the CTA's staging area once and thread (channel, angle) adds the
  //          S sub-frequencies of its channel in index order.
  // The staging memory is DYNAMIC shared memory, present only for FANOUT / BP.
  extern __shared__ __align__(16) double stage[];   // FANOUT: (kRteThreads / 2) * NA + 2 doubles; BP: NA * blockDim.x
  const int lane = threadIdx.x & 31;
  double* wst = stage + ((threadIdx.x - lane) / 2) * NA;
  const long long widx0 = idx_raw - lane;                                  // first pair of this warp
  const long long obase = idx * (long long)A + a0;
  auto emit = [&](double* __restrict__ field, auto value) {
    if (!field) return;
    if constexpr (BP) {
      const int T = blockDim.x, S = g.bpS, C = g.bpC;
      __syncthreads();            // the previous field's readers are done with the staging area
#pragma unroll
      for (int a = 0; a < NA; ++a) stage[a * T + threadIdx.x] = value(a);   // angle-major: conflict-free
      __syncthreads();
      const int c0 = bp_c - (int)(threadIdx.x / S);                        // first channel of this CTA
      for (int r = threadIdx.x; r < g.bpCpb * NA; r += T) {
        const int cl = r / NA, a = r - cl * NA, c = c0 + cl;
        if (c < C && a0 + a < A) {
          const double* src = stage + a * T + cl * S;
          const double* w = g.bpw ? g.bpw + (long long)c * S : nullptr;
          double acc = 0.0;
          for (int sidx = 0; sidx < S; ++sidx) acc = fma(w ? w[sidx] : 1.0 / S, src[sidx], acc);
          field[((long long)p * C + c) * A + a0 + a] = acc;
        }
      }
    } else if constexpr (!FANOUT) {
      if (active) {
#pragma unroll
        for (int a = 0; a < NA; ++a)
          if (a0 + a < A) field[obase + a] = value(a);
      }
    } else {
      const bool coalesced = (A == NA) && ((reinterpret_cast<unsigned long long>(field) & 15ull) == 0);
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        // the 16 threads of half h put their NA values into the warp's slice
        if ((lane >> 4) == h) {
#pragma unroll
          for (int a = 0; a < NA; ++a) wst[(lane & 15) * NA + a] = value(a);
        }
        __syncwarp();
        const long long hidx0 = widx0 + 16 * h;                              // first pair of this half
        const long long hleft = npf - hidx0;
        const int nval = (int)(hleft >= 16 ? 16 : (hleft > 0 ? hleft : 0)) * NA;   // valid doubles
        if (coalesced) {
          double* dst = field + hidx0 * NA;
          const int npeer = (fo.mode == MWRT_OUT_PEERS) ? fo.n : 1;
#pragma unroll 1
          for (int r = 0; r < npeer; ++r) {
            double* d = (fo.mode == MWRT_OUT_PEERS)
                            ? reinterpret_cast<double*>(reinterpret_cast<char*>(dst) + fo.off[r]) : dst;
#pragma unroll 1
            for (int gq = 2 * lane; gq < nval; gq += 64) {
              const double v0 = wst[gq], v1 = wst[gq + 1];     // slice is padded by one pair
              if (gq + 1 < nval) {
                if (fo.mode == MWRT_OUT_MULTICAST) out_store2_mc(d + gq, v0, v1);
                else *reinterpret_cast<double2*>(d + gq) = make_double2(v0, v1);
              } else {
                if (fo.mode == MWRT_OUT_MULTICAST) out_store_mc(d + gq, v0);
                else d[gq] = v0;
              }
            }
          }
        } else if (active && (lane >> 4) == h) {
#pragma unroll 1
          for (int a = 0; a < NA; ++a)
            if (a0 + a < A) out_store(field + obase + a, wst[(lane & 15) * NA + a], fo);
        }
        __syncwarp();
      }
    }
  };
  emit(tbtotal, v_tot);
  emit(tbatm, v_atm);
  emit(tmr, v_mr);
  emit(tauwet, [&](int a) { return bad ? qn : (PLANE ? sw[0] * amass[a] : sw[a]); });
  emit(taudry, [&](int a) {
    return bad ? qn : (PLANE ? sd[0] * amass[a] : (CLOUDY ? tauprof[a] - sw[a] - sl[a] - si[a] : tauprof[a] - sw[a]));
  });
  emit(tauliq, [&](int a) { return bad ? qn : (PLANE ? sl[0] * amass[a] : sl[a]); });
  emit(tauice, [&](int a) { return bad ? qn : (PLANE ? si[0] * amass[a] : si[a]); });
}

template <int NA, bool FANOUT>
static int launch_rte_na(const RteArgs& g, bool cloudy, int from_sat, const Fanout& fo, cudaStream_t st) {
  const long long n = (long long)g.P * g.F;
  const bool plane = (g.ds == nullptr);
  const unsigned gy = (unsigned)((g.A + NA - 1) / NA);
  if constexpr (!FANOUT) {
    if (g.bpC > 0) {   // fused band-pass: plane-parallel, plain stores (checked by rte_launch)
      const int ngrp = (g.bpC + g.bpCpb - 1) / g.bpCpb;
      const int threads = ((g.bpCpb * g.bpS + 31) / 32) * 32;
      dim3 grid((unsigned)((long long)g.P * ngrp), gy);
      const size_t dsm = sizeof(double) * (size_t)threads * NA;
#define MWRT_K_BP(C_, S_) k_rte<NA, C_, true, S_, false, true><<<grid, threads, dsm, st>>>(g, fo)
      if (cloudy && from_sat) MWRT_K_BP(true, true);
      else if (cloudy) MWRT_K_BP(true, false);
      else if (from_sat) MWRT_K_BP(false, true);
      else MWRT_K_BP(false, false);
#undef MWRT_K_BP
      MWRT_CUDA_OK(cudaGetLastError());
      return 0;
    }
  }
  dim3 grid((unsigned)((n + kRteThreads - 1) / kRteThreads), gy);
  const size_t dsm = FANOUT ? sizeof(double) * ((kRteThreads / 2) * NA + 2) : 0;
#define MWRT_K_LAUNCH(C_, P_)                                                                    \
  do {                                                                                           \
    if (from_sat) k_rte<NA, C_, P_, true, FANOUT, false><<<grid, kRteThreads, dsm, st>>>(g, fo); \
    else k_rte<NA, C_, P_, false, FANOUT, false><<<grid, kRteThreads, dsm, st>>>(g, fo);         \
  } while (0)
  if (cloudy && plane) MWRT_K_LAUNCH(true, true);
  else if (cloudy) MWRT_K_LAUNCH(true, false);
  else if (plane) MWRT_K_LAUNCH(false, true);
  else MWRT_K_LAUNCH(false, false);
#undef MWRT_K_LAUNCH
  MWRT_CUDA_OK(cudaGetLastError());
  return 0;
}

// g.ds != null: general path (P,A,L).  g.ds == null: plane-parallel, z (P,L) + amass (A,) device arrays.
// cloudy = liquid/ice terms present: (aliq, aice) arrays, or (denliq, denice) for inline evaluation.
template <bool FANOUT>
static int rte_launch_t(const RteArgs& g, int from_sat, const Fanout& fo, cudaStream_t st) {
  const bool cloudy = (g.aliq != nullptr && g.aice != nullptr) || (g.denliq != nullptr && g.denice != nullptr);
  const int A = g.A;
  int na;
  if (A <= 1) na = 1;
  else if (A <= 2) na = 2;
  else if (A <= 4) na = 4;
  else if (A <= 5) na = 5;
  else if (A <= 8) na = 8;
  else if (A <= 10) na = 10;
  else na = (((A + 9) / 10) * 10 <= ((A + 7) / 8) * 8) ? 10 : 8;
  switch (na) {
    case 1: return launch_rte_na<1, FANOUT>(g, cloudy, from_sat, fo, st);
    case 2: return launch_rte_na<2, FANOUT>(g, cloudy, from_sat, fo, st);
    case 4: return launch_rte_na<4, FANOUT>(g, cloudy, from_sat, fo, st);
    case 5: return launch_rte_na<5, FANOUT>(g, cloudy, from_sat, fo, st);
    case 8: return launch_rte_na<8, FANOUT>(g, cloudy, from_sat, fo, st);
    default: return launch_rte_na<10, FANOUT>(g, cloudy, from_sat, fo, st);
  }
}

}  // namespace mwrt
