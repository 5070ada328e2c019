// mwrt_api.cu -- the extern "C" surface of libmwrt.so (include/mwrt.h): model handles, stage entry
// points, and the plan that chains the kernels for a batch of profiles (device or host buffers).
//
// Boundary it replaces: TbCloudRTE(z,p,t,rh,frq,ang) / init_absmdl / satellite / execute() as called at
// /root/reference/python_src/proc/PyRTlib_processing.py:123-127 -- see INTEGRATION.md.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "mwrt_common.cuh"
#include "mwrt_rte_args.h"
#include <nvtx3/nvToolsExt.h>
#include <stdio.h>
#include <time.h>
#include <stdlib.h>

namespace mwrt {

static thread_local std::string g_err;

int fail_cuda(cudaError_t e, const char* what) {
  g_err = std::string("CUDA error ") + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ") in " + what;
  return (int)e;
}
int fail_arg(int code, const char* what) {
  g_err = what;
  return code;
}

// launchers defined in mwrt_absorption.cu / mwrt_rte.cu
int absorption_launch(const ModelDev& m, int P, int L, int F, const double* p, const double* tk,
                      const double* e, int e_is_rh, double* e_out, double* rho_out, const double* frq,
                      double* awet, double* adry, int* status, int f32ls, AbsBlocks* cache, cudaStream_t st);
int cloud_abs_launch(const ModelDev& m, int P, int L, int F, const double* tk, const double* denliq,
                     const double* denice, const double* frq, double* aliq, double* aice, cudaStream_t st);
int vapor_launch(long long n, const double* tk, const double* rh, double* e, double* rho, cudaStream_t st);
int rte_launch(const RteArgs& g, int from_sat, const Fanout& fo, cudaStream_t st);
int tmrcld_launch(const RteArgs& g, const int* ibase, const int* itop, double* out, cudaStream_t st);
int slant_path_launch(int P, int L, int A, const double* z, const double* p, const double* tk,
                      const double* e, const double* angles, int ray_tracing, double* ds, int* status,
                      cudaStream_t st);
int bandpass_launch(int P, int C, int S, int A, const double* in, const double* w, double* out,
                    const Fanout& fo, cudaStream_t st);
int push_launch(const void* src, long long bytes, void* mc_dst, const long long* peer_dst, int npeer,
                int nblocks, cudaStream_t st);
int dfma_peak(double* flops, double* mhz);
int stats_launch(int T, long long M, const double* values, const double* refs, double* bias,
                 double* sd, double* rmse, double* nvalid, cudaStream_t st);
int thin_mean_launch(int S, int N, int M, const double* h, const double* t, const double* p,
                     const double* rh, const int* inds, const int* n_inds, const double* z_offset,
                     double* zo, double* to, double* po, double* ro, cudaStream_t st);
int assemble_launch(int S, int M, int K, int n_levels, double min_p, const double* zt, const double* tt,
                    const double* pt, const double* rt, const int* n_inds, const double* cz_km,
                    const double* cp, const double* ct, const double* cgkg, const double* crh,
                    double* o_p, double* o_t, double* o_ppmv, double* o_h2o, double* o_z, double* o_rh,
                    int* ok, cudaStream_t st);

int cloud_features_launch(int S, int L, const double* p, const double* t, const double* z, const double* rh,
                          double* lwc_m3, double* lwc_kg, double* iwc_m3, double* iwc_kg, double* lwp,
                          double* iwp, cudaStream_t st);
int math_selftest_launch(int op, long long n, const double* x, double* y, cudaStream_t st);

struct mwrt_builtin_model {
  const char* name;
  mwrt_model_desc desc;
};
#include "mwrt_builtin_models.inc"

}  // namespace mwrt

using namespace mwrt;

struct mwrt_model {
  ModelDev dev;
  double* d_tables;
  double* d_rows;
};

constexpr int kSlots = 4;   // workspace slots = compute streams of the host pipeline

struct mwrt_plan {
  // every member carries a default initialiser: `new mwrt_plan()` yields a plan that
  // mwrt_plan_destroy can always take apart, however far mwrt_plan_create got
  ModelDev model{};
  int L = 0, F = 0, A = 0, flags = 0, chunk = 0;
  double emis = 1.0;
  int C = 0, S = 0;            // band-pass (C == 0: off)
  int bp_cpb = 0;              // > 0: the band-pass is reduced inside kernel 3, this many channels per CTA
  int f32ls = 0;               // fp32 line-shape mode of kernel 1 (mwrt_plan_set_precision)
  AbsBlocks abs_blocks{};      // kernel 1's per-frequency-block scratch, built on first use
  double* d_frq = nullptr;
  double* d_angles = nullptr;
  double* d_amass = nullptr;   // (A,) 1/sin(elevation), computed on the host like PyRTlib does
  double* d_bpw = nullptr;     // (C,S) weights or null (uniform)
  // workspace for `cap` profiles, in `nslots` (1 or 2) independent slots: with two slots kernel 3 of
  // chunk c runs on `s_aux` while kernel 1 of chunk c+1 runs on the caller's stream
  int cap = 0, nslots = 0, split = 1;
  Fanout fanout{MWRT_OUT_PLAIN, 0, {0}};   // how kernel 3 stores the (P,F,A) results
  struct Slot {
    double *awet = nullptr, *adry = nullptr, *aliq = nullptr, *aice = nullptr, *ds = nullptr, *e = nullptr;
    // monochromatic RTE outputs, only for the UNFUSED band-pass fallback (sub-bands wider than one CTA,
    // ray tracing, or fan-out stores); the normal band-pass is reduced inside kernel 3
    double* mono[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_abs = nullptr, ev_rte = nullptr;
    bool used = false;
  } ws[kSlots];
  cudaStream_t s_aux = nullptr;
  // host-path resources (lazily created)
  // Device staging holds the WHOLE batch (inputs, outputs, status: 100 MB for 10^4 profiles), so
  // neither copy engine ever waits for a buffer to be recycled; only the (P,F,L) tau_layer field,
  // which can be tens of GB, is staged per chunk in two slots.
  bool host_ready = false;
  cudaStream_t s_h2d = nullptr, s_comp = nullptr, s_d2h = nullptr;
  cudaStream_t s_cmp[kSlots] = {nullptr, nullptr, nullptr, nullptr};   // s_cmp[0] = s_comp, s_cmp[1] = s_aux, others own
  std::vector<cudaEvent_t> ev_h2d, ev_comp;   // one per chunk
  cudaEvent_t ev_tl[2] = {nullptr, nullptr};  // tau_layer slot copied back
  int hcap = 0, hcap_tl = 0;
  size_t mem_pitch = 0;        // cudaDeviceProp::memPitch: upper bound of a 2-D copy's pitches
  double* h_in = nullptr;      // 6 x hcap x L, field-major
  double* h_out = nullptr;     // 7 x hcap x Fo x A, field-major
  double* h_tl[2] = {nullptr, nullptr};
  int* h_status = nullptr;
};
static_assert(kSlots == 4, "s_cmp initialiser lists four streams");

extern "C" {

int mwrt_version(void) { return MWRT_VERSION; }
const char* mwrt_last_error(void) { return g_err.c_str(); }
int mwrt_set_device(int device) {
  MWRT_CUDA_OK(cudaSetDevice(device));
  return 0;
}
int mwrt_get_device(int* device) {
  if (!device) return fail_arg(MWRT_E_ARG, "mwrt_get_device: null argument");
  MWRT_CUDA_OK(cudaGetDevice(device));
  return 0;
}

int mwrt_builtin_model_count(void) { return k_n_builtin_models; }
const char* mwrt_builtin_model_name(int i) {
  if (i < 0 || i >= k_n_builtin_models) return nullptr;
  return k_builtin_models[i].name;
}
int mwrt_builtin_model_desc(const char* name, mwrt_model_desc* out) {
  if (!name || !out) return fail_arg(MWRT_E_ARG, "mwrt_builtin_model_desc: null argument");
  for (int i = 0; i < k_n_builtin_models; ++i)
    if (strcmp(name, k_builtin_models[i].name) == 0) {
      *out = k_builtin_models[i].desc;
      return 0;
    }
  g_err = std::string("absorption model '") + name + "' is not implemented";
  return MWRT_E_MODEL;
}

int mwrt_model_create_custom(const mwrt_model_desc* d, mwrt_model** out) {
  if (!d || !out || !d->h2o_lines || !d->o2_lines)
    return fail_arg(MWRT_E_ARG, "mwrt_model_create_custom: null argument");
  if (d->n_h2o < 1 || d->n_h2o > MWRT_MAX_H2O_LINES || d->n_o2 < 1 || d->n_o2 > MWRT_MAX_O2_LINES)
    return fail_arg(MWRT_E_ARG, "mwrt_model_create_custom: line count out of range");
  if (d->o2_mixing != 1 && d->o2_mixing != 2)
    return fail_arg(MWRT_E_ARG, "mwrt_model_create_custom: o2_mixing must be 1 or 2");
  mwrt_model* m = new mwrt_model();
  const int nh = d->n_h2o * MWRT_H2O_NCOL, no = d->n_o2 * MWRT_O2_NCOL;
  std::vector<double> host((size_t)nh + no);
  memcpy(host.data(), d->h2o_lines, sizeof(double) * nh);
  memcpy(host.data() + nh, d->o2_lines, sizeof(double) * no);
  cudaError_t e = cudaMalloc(&m->d_tables, sizeof(double) * host.size());
  if (e == cudaSuccess)
    e = cudaMemcpy(m->d_tables, host.data(), sizeof(double) * host.size(), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    delete m;
    return fail_cuda(e, "mwrt_model_create_custom");
  }
  {   // line-major rows for kernel 1's staging (see ModelDev)
    const int nhl = d->n_h2o, nol = d->n_o2, nop = (nol + 1) & ~1;
    std::vector<double> rows((size_t)nhl * kHS + (size_t)nop * kOS, 0.0);
    for (int i = 0; i < nhl; ++i) {
      for (int c = 0; c < MWRT_H2O_NCOL; ++c) rows[(size_t)i * kHS + c] = d->h2o_lines[(size_t)c * nhl + i];
      const double fl = d->h2o_lines[(size_t)MWRT_H2O_FL * nhl + i];
      rows[(size_t)i * kHS + kH_IFL2] = 1.0 / (fl * fl);
    }
    double* orow = rows.data() + (size_t)nhl * kHS;
    for (int i = 0; i < nop; ++i) {
      const int src = i < nol ? i : nol - 1;
      for (int c = 0; c < MWRT_O2_NCOL; ++c) orow[(size_t)i * kOS + c] = d->o2_lines[(size_t)c * nol + src];
      const double fk = d->o2_lines[(size_t)MWRT_O2_F * nol + src];
      orow[(size_t)i * kOS + kO_IFK2] = 1.0 / (fk * fk);
      if (i >= nol) orow[(size_t)i * kOS + MWRT_O2_S300] = 0.0;   // dummy line: no strength
    }
    e = cudaMalloc(&m->d_rows, sizeof(double) * rows.size());
    if (e == cudaSuccess)
      e = cudaMemcpy(m->d_rows, rows.data(), sizeof(double) * rows.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      cudaFree(m->d_tables);
      cudaFree(m->d_rows);
      delete m;
      return fail_cuda(e, "mwrt_model_create_custom");
    }
    m->dev.rows = m->d_rows;
    m->dev.n_rows = (int)rows.size();
  }
  ModelDev& v = m->dev;
  v.n_h2o = d->n_h2o; v.n_o2 = d->n_o2;
  v.h2o_sd = d->h2o_sd; v.o2_mixing = d->o2_mixing; v.o2_dens118 = d->o2_dens118;
  v.o2_sd118 = d->o2_sd118; v.n2_fdep = d->n2_fdep; v.liq_model = d->liq_model;
  v.h2o_reftline = d->h2o_reftline; v.h2o_reftcon = d->h2o_reftcon; v.h2o_cf = d->h2o_cf;
  v.h2o_xcf = d->h2o_xcf; v.h2o_cs = d->h2o_cs; v.h2o_xcs = d->h2o_xcs;
  v.h2o_pvap_div = d->h2o_pvap_div; v.h2o_den_const = d->h2o_den_const;
  v.h2o_line_const = d->h2o_line_const;
  v.o2_x = d->o2_x; v.o2_wb300 = d->o2_wb300; v.o2_pvap_div = d->o2_pvap_div;
  v.o2_h2o_broad = d->o2_h2o_broad; v.o2_nonres = d->o2_nonres;
  v.o2_abs_const = d->o2_abs_num / d->o2_abs_den; v.o2_iso = d->o2_iso;
  v.n2_l = d->n2_l; v.n2_m = d->n2_m; v.n2_n = d->n2_n;
  v.tables = m->d_tables;
  v.n_table = nh + no;
  *out = m;
  return 0;
}

int mwrt_model_create(const char* name, mwrt_model** out) {
  mwrt_model_desc d;
  int rc = mwrt_builtin_model_desc(name, &d);
  if (rc) return rc;
  return mwrt_model_create_custom(&d, out);
}

int mwrt_model_destroy(mwrt_model* m) {
  if (!m) return 0;
  cudaFree(m->d_tables);
  cudaFree(m->d_rows);
  delete m;
  return 0;
}

// ------------------------------------------------------------------------------------ stage API
int mwrt_vapor_f64(int64_t n, const double* tk, const double* rh, double* e, double* rho, void* stream) {
  if (n <= 0 || !tk || !rh || !e) return fail_arg(MWRT_E_ARG, "mwrt_vapor_f64: bad argument");
  return vapor_launch(n, tk, rh, e, rho, (cudaStream_t)stream);
}

static int absorption_stage(const mwrt_model* m, int P, int L, int F, const double* p, const double* tk,
                            const double* e, const double* denliq, const double* denice, const double* frq,
                            double* awet, double* adry, double* aliq, double* aice, int32_t* status,
                            int f32ls, void* stream) {
  if (!m || P <= 0 || L <= 0 || F <= 0 || !p || !tk || !e || !frq || !awet || !adry)
    return fail_arg(MWRT_E_ARG, "mwrt_absorption: bad argument");
  int rc = absorption_launch(m->dev, P, L, F, p, tk, e, 0, nullptr, nullptr, frq, awet, adry, status, f32ls,
                             /*cache=*/nullptr, (cudaStream_t)stream);
  if (rc) return rc;
  if (aliq || aice)   // RTEquation.cloudy_absorption, elementwise (the fused path evaluates it inside kernel 3)
    rc = cloud_abs_launch(m->dev, P, L, F, tk, denliq, denice, frq, aliq, aice, (cudaStream_t)stream);
  return rc;
}

int mwrt_absorption_f64(const mwrt_model* m, int P, int L, int F, const double* p, const double* tk,
                        const double* e, const double* denliq, const double* denice, const double* frq,
                        double* awet, double* adry, double* aliq, double* aice, int32_t* status,
                        void* stream) {
  return absorption_stage(m, P, L, F, p, tk, e, denliq, denice, frq, awet, adry, aliq, aice, status, 0, stream);
}

int mwrt_absorption_f32ls(const mwrt_model* m, int P, int L, int F, const double* p, const double* tk,
                          const double* e, const double* denliq, const double* denice, const double* frq,
                          double* awet, double* adry, double* aliq, double* aice, int32_t* status,
                          void* stream) {
  return absorption_stage(m, P, L, F, p, tk, e, denliq, denice, frq, awet, adry, aliq, aice, status, 1, stream);
}

int mwrt_slant_path_f64(int P, int L, int A, const double* z, const double* p, const double* tk,
                        const double* e, const double* angles, int ray_tracing, double* ds,
                        int32_t* status, void* stream) {
  if (P <= 0 || L <= 0 || A <= 0 || !z || !angles || !ds)
    return fail_arg(MWRT_E_ARG, "mwrt_slant_path_f64: bad argument");
  if (ray_tracing && (!p || !tk || !e))
    return fail_arg(MWRT_E_ARG, "mwrt_slant_path_f64: ray tracing needs p, tk, e");
  return slant_path_launch(P, L, A, z, p, tk, e, angles, ray_tracing, ds, status, (cudaStream_t)stream);
}

int mwrt_rte_f64(int P, int L, int F, int A, const double* tk, const double* awet, const double* adry,
                 const double* aliq, const double* aice, const double* ds, const double* frq,
                 int from_sat, double emissivity, double* tbtotal, double* tbatm, double* tmr,
                 double* tauwet, double* taudry, double* tauliq, double* tauice, double* tau_layer,
                 int tau_layer_angle, int32_t* status, void* stream) {
  if (P <= 0 || L <= 1 || F <= 0 || A <= 0 || !tk || !awet || !adry || !ds || !frq)
    return fail_arg(MWRT_E_ARG, "mwrt_rte_f64: bad argument");
  if ((aliq == nullptr) != (aice == nullptr))
    return fail_arg(MWRT_E_ARG, "mwrt_rte_f64: aliq and aice must be given together");
  if (tau_layer && (tau_layer_angle < 0 || tau_layer_angle >= A))
    return fail_arg(MWRT_E_ARG, "mwrt_rte_f64: tau_layer_angle out of range");
  RteArgs g{};
  g.P = P; g.L = L; g.F = F; g.A = A;
  g.tk = tk; g.awet = awet; g.adry = adry; g.aliq = aliq; g.aice = aice;
  g.ds = ds; g.frq = frq; g.emis = emissivity;
  g.tbtotal = tbtotal; g.tbatm = tbatm; g.tmr = tmr; g.tauwet = tauwet; g.taudry = taudry;
  g.tauliq = tauliq; g.tauice = tauice; g.tau_layer = tau_layer; g.tl_angle = tau_layer_angle;
  g.status = status;
  return rte_launch(g, from_sat, Fanout{MWRT_OUT_PLAIN, 0, {0}}, (cudaStream_t)stream);
}

int mwrt_bandpass_f64(int P, int C, int S, int A, const double* in, const double* w, double* out,
                      void* stream) {
  if (P <= 0 || C <= 0 || S <= 0 || A <= 0 || !in || !out)
    return fail_arg(MWRT_E_ARG, "mwrt_bandpass_f64: bad argument");
  return bandpass_launch(P, C, S, A, in, w, out, Fanout{MWRT_OUT_PLAIN, 0, {0}}, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------ plan
// is the band-pass of this plan reduced inside kernel 3?  (else: monochromatic fields + k_bandpass)
static bool bp_fused(const mwrt_plan* pl) {
  return pl->C > 0 && pl->bp_cpb > 0 && !(pl->flags & MWRT_F_RAY_TRACING) && pl->fanout.mode == MWRT_OUT_PLAIN;
}

static void plan_free_ws(mwrt_plan* pl) {
  for (int k = 0; k < kSlots; ++k) {
    mwrt_plan::Slot& w = pl->ws[k];
    cudaFree(w.awet); cudaFree(w.adry); cudaFree(w.ds); cudaFree(w.e);
    for (int i = 0; i < 7; ++i) cudaFree(w.mono[i]);
    w.awet = w.adry = w.ds = w.e = nullptr;
    for (int i = 0; i < 7; ++i) w.mono[i] = nullptr;
    w.used = false;
  }
  pl->cap = 0;
  pl->nslots = 0;
}

// Workspace of one slot: awet, adry (cap,F,L) -- cloud liquid/ice absorption is evaluated inside
// kernel 3 and has no array --, e (cap,L), ds (cap,A,L) for ray tracing; the monochromatic result
// fields only for the unfused band-pass fallback.
static int plan_reserve(mwrt_plan* pl, int cap, int nslots) {
  const bool need_mono = pl->C > 0 && !bp_fused(pl);
  const bool have_mono = pl->nslots > 0 && pl->ws[0].mono[0] != nullptr;
  if (cap <= pl->cap && nslots <= pl->nslots && (!need_mono || have_mono)) return 0;
  cap = std::max(cap, pl->cap);
  nslots = std::max(nslots, pl->nslots);
  plan_free_ws(pl);
  const size_t pfl = sizeof(double) * (size_t)cap * pl->F * pl->L;
  for (int k = 0; k < nslots; ++k) {
    mwrt_plan::Slot& w = pl->ws[k];
    MWRT_CUDA_OK(cudaMalloc(&w.awet, pfl));
    MWRT_CUDA_OK(cudaMalloc(&w.adry, pfl));
    if (pl->flags & MWRT_F_RAY_TRACING)
      MWRT_CUDA_OK(cudaMalloc(&w.ds, sizeof(double) * (size_t)cap * pl->A * pl->L));
    MWRT_CUDA_OK(cudaMalloc(&w.e, sizeof(double) * (size_t)cap * pl->L));
    if (need_mono)
      for (int i = 0; i < 7; ++i)
        MWRT_CUDA_OK(cudaMalloc(&w.mono[i], sizeof(double) * (size_t)cap * pl->F * pl->A));
  }
  pl->cap = cap;
  pl->nslots = nslots;
  return 0;
}

static int default_chunk(int L, int F) {
  // bound the two (chunk,F,L) absorption arrays to ~1 GiB; they are streamed once (write, then read)
  // and cost < 5 % of the FP64 time even when they spill out of the 126 MB L2
  const double per_profile = 8.0 * F * L * 2;
  double c = (1024.0 * 1024.0 * 1024.0) / per_profile;
  if (c < 64) c = 64;
  if (c > (1 << 20)) c = (1 << 20);
  return (int)c;
}

int mwrt_plan_create(const mwrt_model* m, int L, int F, int A, const double* frq, const double* angles,
                     int flags, int chunk_profiles, mwrt_plan** out) {
  if (!m || !frq || !angles || !out || L < 2 || F < 1 || A < 1)
    return fail_arg(MWRT_E_ARG, "mwrt_plan_create: bad argument");
  mwrt_plan* pl = new mwrt_plan();   // value-initialised: see the member initialisers
  pl->model = m->dev;
  pl->L = L; pl->F = F; pl->A = A; pl->flags = flags;
  pl->chunk = chunk_profiles > 0 ? chunk_profiles : default_chunk(L, F);
  cudaError_t e = cudaStreamCreateWithFlags(&pl->s_aux, cudaStreamNonBlocking);
  for (int k = 0; k < 2 && e == cudaSuccess; ++k) {
    e = cudaEventCreateWithFlags(&pl->ws[k].ev_abs, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&pl->ws[k].ev_rte, cudaEventDisableTiming);
  }
  if (e == cudaSuccess) e = cudaMalloc(&pl->d_frq, sizeof(double) * F);
  if (e == cudaSuccess) e = cudaMalloc(&pl->d_angles, sizeof(double) * A);
  if (e == cudaSuccess) e = cudaMemcpy(pl->d_frq, frq, sizeof(double) * F, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(pl->d_angles, angles, sizeof(double) * A, cudaMemcpyHostToDevice);
  std::vector<double> amass(A);
  for (int a = 0; a < A; ++a) amass[a] = 1.0 / sin(angles[a] * 3.141592653589793 / 180.0);
  if (e == cudaSuccess) e = cudaMalloc(&pl->d_amass, sizeof(double) * A);
  if (e == cudaSuccess) e = cudaMemcpy(pl->d_amass, amass.data(), sizeof(double) * A, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    int dev = 0;
    cudaDeviceProp prop;
    e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, dev);
    if (e == cudaSuccess) pl->mem_pitch = prop.memPitch;
  }
  if (e != cudaSuccess) {
    mwrt_plan_destroy(pl);   // releases whatever exists: stream, events, device arrays
    return fail_cuda(e, "mwrt_plan_create");
  }
  *out = pl;
  return 0;
}

static void plan_free_host(mwrt_plan* pl) {
  if (!pl->host_ready) return;
  cudaFree(pl->h_in);
  cudaFree(pl->h_out);
  cudaFree(pl->h_status);
  for (int s = 0; s < 2; ++s) {
    cudaFree(pl->h_tl[s]);
    pl->h_tl[s] = nullptr;
    if (pl->ev_tl[s]) cudaEventDestroy(pl->ev_tl[s]);
    pl->ev_tl[s] = nullptr;
  }
  for (cudaEvent_t e : pl->ev_h2d) cudaEventDestroy(e);
  for (cudaEvent_t e : pl->ev_comp) cudaEventDestroy(e);
  pl->ev_h2d.clear();
  pl->ev_comp.clear();
  if (pl->s_h2d) cudaStreamDestroy(pl->s_h2d);
  if (pl->s_comp) cudaStreamDestroy(pl->s_comp);
  if (pl->s_d2h) cudaStreamDestroy(pl->s_d2h);
  for (int k = 2; k < kSlots; ++k)
    if (pl->s_cmp[k]) cudaStreamDestroy(pl->s_cmp[k]);
  pl->s_h2d = pl->s_comp = pl->s_d2h = nullptr;
  for (int k = 0; k < kSlots; ++k) pl->s_cmp[k] = nullptr;
  pl->h_in = pl->h_out = nullptr;
  pl->h_status = nullptr;
  pl->host_ready = false;
  pl->hcap = pl->hcap_tl = 0;
}

int mwrt_plan_destroy(mwrt_plan* pl) {
  if (!pl) return 0;
  plan_free_host(pl);
  plan_free_ws(pl);
  for (int k = 0; k < 2; ++k) {
    if (pl->ws[k].ev_abs) cudaEventDestroy(pl->ws[k].ev_abs);
    if (pl->ws[k].ev_rte) cudaEventDestroy(pl->ws[k].ev_rte);
  }
  if (pl->s_aux) cudaStreamDestroy(pl->s_aux);
  cudaFree(pl->d_frq); cudaFree(pl->d_angles); cudaFree(pl->d_amass); cudaFree(pl->d_bpw);
  cudaFree(pl->abs_blocks.buf);
  delete pl;
  return 0;
}

int mwrt_plan_set_emissivity(mwrt_plan* pl, double es) {
  if (!pl) return fail_arg(MWRT_E_ARG, "mwrt_plan_set_emissivity: null plan");
  pl->emis = es;
  return 0;
}

int mwrt_plan_absorption(mwrt_plan* pl, int P, const double* p, const double* tk, const double* e,
                         double* awet, double* adry, int32_t* status, void* stream) {
  if (!pl || !p || !tk || !e || !awet || !adry || P < 0)
    return fail_arg(MWRT_E_ARG, "mwrt_plan_absorption: bad argument");
  if (P == 0) return 0;
  return absorption_launch(pl->model, P, pl->L, pl->F, p, tk, e, /*e_is_rh=*/0, nullptr, nullptr, pl->d_frq,
                           awet, adry, status, pl->f32ls, &pl->abs_blocks, (cudaStream_t)stream);
}

int mwrt_plan_set_precision(mwrt_plan* pl, int precision) {
  if (!pl || (precision != MWRT_PREC_F64 && precision != MWRT_PREC_F32_LINESHAPE))
    return fail_arg(MWRT_E_ARG, "mwrt_plan_set_precision: bad argument");
  pl->f32ls = precision == MWRT_PREC_F32_LINESHAPE;
  return 0;
}

int mwrt_plan_set_output_fanout(mwrt_plan* pl, int mode, int npeer, const int64_t* peer_byte_offsets) {
  if (!pl || mode < MWRT_OUT_PLAIN || mode > MWRT_OUT_PEERS)
    return fail_arg(MWRT_E_ARG, "mwrt_plan_set_output_fanout: bad argument");
  Fanout f{mode, 0, {0}};
  if (mode == MWRT_OUT_PEERS) {
    if (npeer < 1 || npeer > MWRT_MAX_PEERS || !peer_byte_offsets)
      return fail_arg(MWRT_E_ARG, "mwrt_plan_set_output_fanout: 1..MWRT_MAX_PEERS byte offsets required");
    f.n = npeer;
    for (int r = 0; r < npeer; ++r) f.off[r] = (long long)peer_byte_offsets[r];
  }
  pl->fanout = f;
  return 0;
}

int mwrt_push_async(void* dst, const void* src, int64_t bytes, void* stream) {
  if (!dst || !src || bytes <= 0) return fail_arg(MWRT_E_ARG, "mwrt_push_async: bad argument");
  MWRT_CUDA_OK(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}

int mwrt_push_kernel(const void* src, int64_t bytes, void* mc_dst, const int64_t* peer_dst, int npeer,
                     int nblocks, void* stream) {
  if (!src || bytes <= 0 || (!mc_dst && !peer_dst)) return fail_arg(MWRT_E_ARG, "mwrt_push_kernel: bad argument");
  static_assert(sizeof(long long) == sizeof(int64_t), "peer addresses are passed as 64-bit integers");
  return push_launch(src, (long long)bytes, mc_dst, reinterpret_cast<const long long*>(peer_dst), npeer, nblocks,
                     (cudaStream_t)stream);
}

int mwrt_plan_set_split(mwrt_plan* pl, int pieces) {
  if (!pl || pieces < 1 || pieces > 64) return fail_arg(MWRT_E_ARG, "mwrt_plan_set_split: bad argument");
  pl->split = pieces;
  return 0;
}

int mwrt_plan_set_bandpass(mwrt_plan* pl, int C, int S, const double* weights) {
  if (!pl) return fail_arg(MWRT_E_ARG, "mwrt_plan_set_bandpass: null plan");
  if (C <= 0 || S <= 0 || (long long)C * S != pl->F)
    return fail_arg(MWRT_E_ARG, "mwrt_plan_set_bandpass: C*S must equal F");
  plan_free_ws(pl);
  plan_free_host(pl);
  cudaFree(pl->d_bpw);
  pl->d_bpw = nullptr;
  if (weights) {
    MWRT_CUDA_OK(cudaMalloc(&pl->d_bpw, sizeof(double) * C * S));
    MWRT_CUDA_OK(cudaMemcpy(pl->d_bpw, weights, sizeof(double) * C * S, cudaMemcpyHostToDevice));
  }
  pl->C = C; pl->S = S;
  pl->bp_cpb = rte_bp_channels_per_cta(C, S);   // 0: a sub-band does not fit one CTA -> unfused fallback
  if (const char* e = getenv("MWRT_BP_CPB")) {   // tuning knob: fewer channels per CTA (more, smaller CTAs per SM)
    const int want = atoi(e);
    if (want >= 1 && want < pl->bp_cpb) pl->bp_cpb = want;
  }
  return 0;
}

int mwrt_plan_launch_count(const mwrt_plan* pl, int P) {
  if (!pl || P <= 0) return 0;
  int chunk = std::min(P, pl->chunk);
  if (pl->split > 1 && P >= 1024 * pl->split) chunk = std::min(chunk, (P + pl->split - 1) / pl->split);
  const int nchunk = (P + chunk - 1) / chunk;
  const int bp = (pl->C > 0 && !bp_fused(pl)) ? 7 : 0;   // unfused fallback: one k_bandpass per field
  return nchunk * (2 + ((pl->flags & MWRT_F_RAY_TRACING) ? 1 : 0) + bp);
}

// one chunk: device pointers already offset to the chunk's first profile.  Kernel 1 (+2) go to
// st_abs, kernel 3 to st_rte; when the two differ the hand-off and the reuse of the workspace slot
// are ordered with the slot's events.
static int forward_chunk(mwrt_plan* pl, int slot, int P, const mwrt_profiles& in,
                         const mwrt_outputs& out, cudaStream_t st_abs, cudaStream_t st_rte) {
  const bool cloudy = pl->flags & MWRT_F_CLOUDY;
  const bool rt = pl->flags & MWRT_F_RAY_TRACING;
  const bool two = st_abs != st_rte;
  const int L = pl->L, F = pl->F, A = pl->A;
  mwrt_plan::Slot& w = pl->ws[slot];
  if (two && w.used) MWRT_CUDA_OK(cudaStreamWaitEvent(st_abs, w.ev_rte, 0));  // slot still being read?
  if (out.status) MWRT_CUDA_OK(cudaMemsetAsync(out.status, 0, sizeof(int) * (size_t)P, st_abs));
  // kernel 1: vapour pressure + gas absorption (e is kept for the refractivity of kernel 2)
  nvtxRangePushA("mwrt:absorption");
  int rc = absorption_launch(pl->model, P, L, F, in.p, in.tk, in.rh, /*e_is_rh=*/1, w.e, nullptr, pl->d_frq,
                             w.awet, w.adry, out.status, pl->f32ls, &pl->abs_blocks, st_abs);
  nvtxRangePop();
  if (rc) return rc;
  // kernel 2: refractive slant path (the plane-parallel default needs no kernel: kernel 3 forms
  // ds = dz / sin(elev) in registers)
  if (rt) {
    nvtxRangePushA("mwrt:slant_path");
    rc = slant_path_launch(P, L, A, in.z, in.p, in.tk, w.e, pl->d_angles, 1, w.ds, out.status, st_abs);
    nvtxRangePop();
    if (rc) return rc;
  }
  if (two) {
    MWRT_CUDA_OK(cudaEventRecord(w.ev_abs, st_abs));
    MWRT_CUDA_OK(cudaStreamWaitEvent(st_rte, w.ev_abs, 0));
  }
  // kernel 3: layer integration + RTE, cloud absorption inline, band-pass reduced in the kernel
  const bool fused = bp_fused(pl);
  double* tl = (pl->flags & MWRT_F_TAU_LAYERS) ? out.tau_layer : nullptr;
  double* dst[7] = {out.tbtotal, out.tbatm, out.tmr, out.tauwet, out.taudry, out.tauliq, out.tauice};
  double* mono[7];
  for (int i = 0; i < 7; ++i) mono[i] = (pl->C > 0 && !fused) ? (dst[i] ? w.mono[i] : nullptr) : dst[i];
  RteArgs g{};
  g.P = P; g.L = L; g.F = F; g.A = A;
  g.tk = in.tk; g.awet = w.awet; g.adry = w.adry;
  g.denliq = cloudy ? in.denliq : nullptr;
  g.denice = cloudy ? in.denice : nullptr;
  g.liq_model = pl->model.liq_model;
  g.ds = rt ? w.ds : nullptr;
  g.z = in.z; g.amass = pl->d_amass; g.frq = pl->d_frq; g.emis = pl->emis;
  g.tbtotal = mono[0]; g.tbatm = mono[1]; g.tmr = mono[2]; g.tauwet = mono[3]; g.taudry = mono[4];
  g.tauliq = mono[5]; g.tauice = mono[6];
  g.tau_layer = tl; g.tl_angle = 0; g.status = out.status;
  if (fused) { g.bpC = pl->C; g.bpS = pl->S; g.bpCpb = pl->bp_cpb; g.bpw = pl->d_bpw; }
  nvtxRangePushA("mwrt:rte");
  rc = rte_launch(g, (pl->flags & MWRT_F_FROM_SAT) ? 1 : 0,
                  (pl->C > 0) ? Fanout{MWRT_OUT_PLAIN, 0, {0}} : pl->fanout, st_rte);
  nvtxRangePop();
  if (rc) return rc;
  if (pl->C > 0 && !fused)
    for (int i = 0; i < 7; ++i)
      if (dst[i]) {
        rc = bandpass_launch(P, pl->C, pl->S, A, mono[i], pl->d_bpw, dst[i], pl->fanout, st_rte);
        if (rc) return rc;
      }
  if (out.tmrcld) {   // rare: upstream's `tmrcld` column (checked by check_io: no band-pass, downwelling)
    g.bpC = 0;
    rc = tmrcld_launch(g, in.cld_ibase, in.cld_itop, out.tmrcld, st_rte);
    if (rc) return rc;
  }
  if (two) {
    MWRT_CUDA_OK(cudaEventRecord(w.ev_rte, st_rte));
    w.used = true;
  }
  return 0;
}

static int check_io(const mwrt_plan* pl, int P, const mwrt_profiles* in, const mwrt_outputs* out,
                    const char* who) {
  if (!pl || !in || !out || P <= 0) return fail_arg(MWRT_E_ARG, who);
  if (!in->z || !in->p || !in->tk || !in->rh) return fail_arg(MWRT_E_ARG, "forward: z/p/tk/rh required");
  if ((pl->flags & MWRT_F_CLOUDY) && (!in->denliq || !in->denice))
    return fail_arg(MWRT_E_ARG, "forward: MWRT_F_CLOUDY needs denliq and denice");
  if (out->tau_layer && !(pl->flags & MWRT_F_TAU_LAYERS))
    return fail_arg(MWRT_E_ARG, "forward: tau_layer needs MWRT_F_TAU_LAYERS");
  if (out->tmrcld) {
    if (!in->cld_ibase || !in->cld_itop)
      return fail_arg(MWRT_E_ARG, "forward: tmrcld needs cld_ibase and cld_itop");
    if (pl->C > 0 || (pl->flags & MWRT_F_FROM_SAT))
      return fail_arg(MWRT_E_ARG, "forward: tmrcld is defined for downwelling, monochromatic runs only");
  }
  return 0;
}

static mwrt_profiles offset_in(const mwrt_profiles& in, size_t off) {
  mwrt_profiles r = in;
  r.z += off; r.p += off; r.tk += off; r.rh += off;
  if (r.denliq) r.denliq += off;
  if (r.denice) r.denice += off;
  return r;
}
static mwrt_profiles offset_in_p(mwrt_profiles r, size_t off_p) {   // the per-profile (P,) members
  if (r.cld_ibase) r.cld_ibase += off_p;
  if (r.cld_itop) r.cld_itop += off_p;
  return r;
}
static mwrt_outputs offset_out(const mwrt_outputs& o, size_t off_pfa, size_t off_pfl, size_t off_p) {
  mwrt_outputs r = o;
  double** f[7] = {&r.tbtotal, &r.tbatm, &r.tmr, &r.tauwet, &r.taudry, &r.tauliq, &r.tauice};
  for (int i = 0; i < 7; ++i)
    if (*f[i]) *f[i] += off_pfa;
  if (r.tau_layer) r.tau_layer += off_pfl;
  if (r.status) r.status += off_p;
  if (r.tmrcld) r.tmrcld += off_pfa;
  return r;
}

int mwrt_plan_forward_device(mwrt_plan* pl, int P, const mwrt_profiles* in, const mwrt_outputs* out,
                             void* stream) {
  int rc = check_io(pl, P, in, out, "mwrt_plan_forward_device: bad argument");
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  // split > 1: cut the batch into pieces and run kernel 3 of piece c on the plan's side stream while
  // kernel 1 of piece c+1 runs on the caller's stream (two workspace slots)
  int chunk = std::min(P, pl->chunk);
  const bool two = pl->split > 1 && P >= 1024 * pl->split;
  if (two) chunk = std::min(chunk, (P + pl->split - 1) / pl->split);
  rc = plan_reserve(pl, chunk, two ? 2 : 1);
  if (rc) return rc;
  const int Fo = pl->C > 0 ? pl->C : pl->F;
  int c = 0;
  for (int p0 = 0; p0 < P; p0 += chunk, ++c) {
    const int n = std::min(chunk, P - p0);
    rc = forward_chunk(pl, two ? (c & 1) : 0, n, offset_in_p(offset_in(*in, (size_t)p0 * pl->L), p0),
                       offset_out(*out, (size_t)p0 * Fo * pl->A, (size_t)p0 * pl->F * pl->L, p0), st,
                       two ? pl->s_aux : st);
    if (rc) return rc;
  }
  if (two) {  // the caller's stream completes only when the last kernel 3 has
    for (int k = 0; k < 2; ++k)
      if (pl->ws[k].used) MWRT_CUDA_OK(cudaStreamWaitEvent(st, pl->ws[k].ev_rte, 0));
  }
  return 0;
}

static int plan_host_init(mwrt_plan* pl, int hcap, int hcap_tl, int nchunks) {
  const bool tl = pl->flags & MWRT_F_TAU_LAYERS;
  if (!(pl->host_ready && hcap <= pl->hcap && (!tl || hcap_tl <= pl->hcap_tl))) {
    plan_free_host(pl);
    MWRT_CUDA_OK(cudaStreamCreateWithFlags(&pl->s_h2d, cudaStreamNonBlocking));
    MWRT_CUDA_OK(cudaStreamCreateWithFlags(&pl->s_comp, cudaStreamNonBlocking));
    MWRT_CUDA_OK(cudaStreamCreateWithFlags(&pl->s_d2h, cudaStreamNonBlocking));
    pl->s_cmp[0] = pl->s_comp;
    pl->s_cmp[1] = pl->s_aux;
    for (int k = 2; k < kSlots; ++k) MWRT_CUDA_OK(cudaStreamCreateWithFlags(&pl->s_cmp[k], cudaStreamNonBlocking));
    pl->host_ready = true;   // from here on plan_free_host releases what exists
    const int Fo = pl->C > 0 ? pl->C : pl->F;
    MWRT_CUDA_OK(cudaMalloc(&pl->h_in, sizeof(double) * 6 * (size_t)hcap * pl->L));
    MWRT_CUDA_OK(cudaMalloc(&pl->h_out, sizeof(double) * 7 * (size_t)hcap * Fo * pl->A));
    MWRT_CUDA_OK(cudaMalloc(&pl->h_status, sizeof(int) * (size_t)hcap));
    for (int s = 0; s < 2; ++s) {
      MWRT_CUDA_OK(cudaEventCreateWithFlags(&pl->ev_tl[s], cudaEventDisableTiming));
      if (tl) MWRT_CUDA_OK(cudaMalloc(&pl->h_tl[s], sizeof(double) * (size_t)hcap_tl * pl->F * pl->L));
    }
    pl->hcap = hcap;
    pl->hcap_tl = hcap_tl;
  }
  while ((int)pl->ev_h2d.size() < nchunks) {
    cudaEvent_t a, b;
    MWRT_CUDA_OK(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
    pl->ev_h2d.push_back(a);
    MWRT_CUDA_OK(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
    pl->ev_comp.push_back(b);
  }
  return 0;
}

// true when the non-null pointers of v[0..n) are equally spaced by `stride` bytes (so one 2-D copy
// moves them all); k0/k1 = first/last non-null index, every index between must be non-null too
static bool equally_spaced(const double* const* v, int n, int& k0, int& k1, ptrdiff_t& stride) {
  k0 = -1; k1 = -1;
  for (int i = 0; i < n; ++i)
    if (v[i]) { if (k0 < 0) k0 = i; k1 = i; }
  if (k0 < 0 || k1 == k0) return false;
  stride = ((const char*)v[k1] - (const char*)v[k0]) / (k1 - k0);
  if (stride <= 0) return false;
  for (int i = k0; i <= k1; ++i)
    if (!v[i] || (const char*)v[i] - (const char*)v[k0] != stride * (i - k0)) return false;
  return true;
}

int mwrt_plan_forward_host(mwrt_plan* pl, int P, const mwrt_profiles* in, const mwrt_outputs* out) {
  int rc = check_io(pl, P, in, out, "mwrt_plan_forward_host: bad argument");
  if (rc) return rc;
  if (pl->fanout.mode != MWRT_OUT_PLAIN)
    return fail_arg(MWRT_E_ARG, "mwrt_plan_forward_host: output fan-out is a device-entry feature");
  if (out->tmrcld)
    return fail_arg(MWRT_E_ARG, "mwrt_plan_forward_host: tmrcld is a device-entry feature");
  const int L = pl->L, F = pl->F, A = pl->A;
  const int Fo = pl->C > 0 ? pl->C : F;
  const bool cloudy = pl->flags & MWRT_F_CLOUDY;
  const bool want_tl = out->tau_layer != nullptr;
  // Chunk schedule.  The step is bound by the device->host copy of the outputs (67 MB against 32 MB
  // of inputs and ~1.2 ms of kernels for config 2), so the copy engine must start early and never
  // idle.  Measured on B200 (tools/host_trace.py): equal chunks of ~1000 profiles on two alternating
  // compute streams beat both fewer, larger chunks (late first copy, long drain) and a geometric
  // ramp (tiny kernels are latency-bound: one kernel-3 thread walks all layers serially) or a taper.
  // MWRT_HOST_PIECES overrides the piece count (tuning aid; 10 pieces of ~1000 is the optimum here).
  static const int env_pieces = [] { const char* e = getenv("MWRT_HOST_PIECES"); return e ? atoi(e) : 0; }();
  const int pieces = env_pieces > 0 ? env_pieces : (pl->split > 1 ? 2 * pl->split : 12);
  int chunk = std::min(pl->chunk, std::max((P + pieces - 1) / pieces, env_pieces > 0 ? 256 : 1024));
  if (const char* ce = getenv("MWRT_HOST_CHUNK")) {   // tuning knob: explicit chunk size in profiles
    const int want = atoi(ce);
    if (want >= 64) chunk = std::min(pl->chunk, want);
  }
  chunk = std::min(chunk, P);
  std::vector<int> sizes;
  for (int left = P; left > 0; left -= sizes.back()) sizes.push_back(std::min(chunk, left));
  const int nchunks = (int)sizes.size();
  static const int nstr = [] { const char* e = getenv("MWRT_HOST_STREAMS"); int v = e ? atoi(e) : 2; return v < 1 ? 1 : (v > kSlots ? kSlots : v); }();
  rc = plan_reserve(pl, chunk, nstr);
  if (rc) return rc;
  rc = plan_host_init(pl, P, chunk, nchunks);
  if (rc) return rc;
  const size_t in_pitch = (size_t)pl->hcap * L, out_pitch = (size_t)pl->hcap * Fo * A;   // doubles
  const double* hin[6] = {in->z, in->p, in->tk, in->rh, cloudy ? in->denliq : nullptr,
                          cloudy ? in->denice : nullptr};
  double* hout[7] = {out->tbtotal, out->tbatm, out->tmr, out->tauwet, out->taudry, out->tauliq, out->tauice};
  // Clear sky: tauliq / tauice are identically zero (NaN for flagged profiles).  They are neither
  // computed nor moved over PCIe (17 % of the output bytes each): the host fills them while the GPU
  // works and patches the flagged profiles once the status words are back.
  double* host_fill[2] = {nullptr, nullptr};
  if (!cloudy) {
    host_fill[0] = hout[5];
    host_fill[1] = hout[6];
    hout[5] = hout[6] = nullptr;
  }
  const int nin = cloudy ? 6 : 4;
  int ik0, ik1, ok0, ok1;
  ptrdiff_t istride = 0, ostride = 0;
  // one 2-D copy per chunk moves all fields when they are equally spaced -- but only while both
  // pitches stay within cudaDeviceProp::memPitch (cudaMemcpy2D rejects larger ones: a dense-spectrum
  // batch can exceed 2 GiB per field); otherwise one cudaMemcpyAsync per field per chunk
  const size_t max_pitch = pl->mem_pitch ? pl->mem_pitch : (size_t)0x7fffffff;
  const char* penv = getenv("MWRT_MAX_PITCH");
  const size_t env_pitch = penv ? (size_t)atoll(penv) : (size_t)0;
  const size_t pitch_cap = env_pitch ? std::min(env_pitch, max_pitch) : max_pitch;   // MWRT_MAX_PITCH: test hook
  bool in_2d = equally_spaced(hin, nin, ik0, ik1, istride) && ik0 == 0 && ik1 == nin - 1;
  bool out_2d = equally_spaced(const_cast<const double* const*>(hout), 7, ok0, ok1, ostride);
  if (in_2d && ((size_t)istride > pitch_cap || sizeof(double) * in_pitch > pitch_cap)) in_2d = false;
  if (out_2d && ((size_t)ostride > pitch_cap || sizeof(double) * out_pitch > pitch_cap)) out_2d = false;

  // MWRT_TRACE=1: timing events per chunk, timeline printed to stderr after the call (tuning aid)
  static const bool trace = [] { const char* e = getenv("MWRT_TRACE"); return e && atoi(e) != 0; }();
  std::vector<cudaEvent_t> tev;
  cudaEvent_t tev0 = nullptr;
  auto mark = [&](cudaStream_t st) {
    if (!trace) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    tev.push_back(e);
  };
  timespec ts0, ts1;
  if (trace) {
    cudaEventCreate(&tev0);
    cudaEventRecord(tev0, pl->s_h2d);
    clock_gettime(CLOCK_MONOTONIC, &ts0);
  }

  // all inputs go up front, chunk by chunk, so the bus is free for results as early as possible
  int p0 = 0;
  for (int c = 0; c < nchunks; p0 += sizes[c], ++c) {
    const size_t off = (size_t)p0 * L, bytes = sizeof(double) * (size_t)sizes[c] * L;
    if (in_2d) {
      MWRT_CUDA_OK(cudaMemcpy2DAsync(pl->h_in + off, sizeof(double) * in_pitch, hin[0] + off, (size_t)istride,
                                     bytes, nin, cudaMemcpyHostToDevice, pl->s_h2d));
    } else {
      for (int i = 0; i < nin; ++i)
        MWRT_CUDA_OK(cudaMemcpyAsync(pl->h_in + i * in_pitch + off, hin[i] + off, bytes,
                                     cudaMemcpyHostToDevice, pl->s_h2d));
    }
    MWRT_CUDA_OK(cudaEventRecord(pl->ev_h2d[c], pl->s_h2d));
    mark(pl->s_h2d);
  }
  p0 = 0;
  for (int c = 0; c < nchunks; p0 += sizes[c], ++c) {
    const int n = sizes[c];
    const int s = c % nstr;
    const int stl = c & 1;   // tau_layer staging slot
    const size_t ioff = (size_t)p0 * L, ooff = (size_t)p0 * Fo * A;
    // Chunk c runs kernel 1 then kernel 3 on compute stream c&1: neighbouring chunks are on different
    // streams, so the tail of one chunk's kernels (partial last wave, the long serial walk of kernel
    // 3) overlaps the next chunk's kernel 1, while chunks c and c-2 -- which share a workspace slot --
    // are ordered by their stream.
    cudaStream_t sc = pl->s_cmp[s];
    MWRT_CUDA_OK(cudaStreamWaitEvent(sc, pl->ev_h2d[c], 0));
    // the tau_layer slot is free once chunk c-2 has been copied back (kernel 3 writes it)
    if (want_tl && c >= 2) MWRT_CUDA_OK(cudaStreamWaitEvent(sc, pl->ev_tl[stl], 0));
    mwrt_profiles din = {pl->h_in + ioff, pl->h_in + in_pitch + ioff, pl->h_in + 2 * in_pitch + ioff,
                         pl->h_in + 3 * in_pitch + ioff, cloudy ? pl->h_in + 4 * in_pitch + ioff : nullptr,
                         cloudy ? pl->h_in + 5 * in_pitch + ioff : nullptr, nullptr, nullptr};
    mwrt_outputs dout{};
    double** df[7] = {&dout.tbtotal, &dout.tbatm, &dout.tmr, &dout.tauwet, &dout.taudry, &dout.tauliq, &dout.tauice};
    for (int i = 0; i < 7; ++i) *df[i] = hout[i] ? pl->h_out + i * out_pitch + ooff : nullptr;
    dout.tau_layer = want_tl ? pl->h_tl[stl] : nullptr;
    dout.status = pl->h_status + p0;
    rc = forward_chunk(pl, s, n, din, dout, sc, sc);
    if (rc) return rc;
    mark(sc);
    mark(sc);
    MWRT_CUDA_OK(cudaEventRecord(pl->ev_comp[c], sc));
    MWRT_CUDA_OK(cudaStreamWaitEvent(pl->s_d2h, pl->ev_comp[c], 0));
    const size_t obytes = sizeof(double) * (size_t)n * Fo * A;
    if (out_2d) {
      MWRT_CUDA_OK(cudaMemcpy2DAsync(hout[ok0] + ooff, (size_t)ostride, pl->h_out + ok0 * out_pitch + ooff,
                                     sizeof(double) * out_pitch, obytes, ok1 - ok0 + 1,
                                     cudaMemcpyDeviceToHost, pl->s_d2h));
    } else {
      for (int i = 0; i < 7; ++i)
        if (hout[i])
          MWRT_CUDA_OK(cudaMemcpyAsync(hout[i] + ooff, pl->h_out + i * out_pitch + ooff, obytes,
                                       cudaMemcpyDeviceToHost, pl->s_d2h));
    }
    if (want_tl) {
      MWRT_CUDA_OK(cudaMemcpyAsync(out->tau_layer + (size_t)p0 * F * L, pl->h_tl[stl],
                                   sizeof(double) * (size_t)n * F * L, cudaMemcpyDeviceToHost, pl->s_d2h));
      MWRT_CUDA_OK(cudaEventRecord(pl->ev_tl[stl], pl->s_d2h));
    }
    mark(pl->s_d2h);
  }
  // status: one small copy at the end (kernel 3 of the last chunk has run once ev_comp fired)
  if (out->status)
    MWRT_CUDA_OK(cudaMemcpyAsync(out->status, pl->h_status, sizeof(int) * (size_t)P, cudaMemcpyDeviceToHost,
                                 pl->s_d2h));
  if (trace) clock_gettime(CLOCK_MONOTONIC, &ts1);
  std::vector<int> st_local;
  int* st_host = out->status;
  if ((host_fill[0] || host_fill[1]) && !st_host) {   // the patch below needs the status words
    st_local.resize(P);
    st_host = st_local.data();
    MWRT_CUDA_OK(cudaMemcpyAsync(st_host, pl->h_status, sizeof(int) * (size_t)P, cudaMemcpyDeviceToHost,
                                 pl->s_d2h));
  }
  for (int k = 0; k < 2; ++k)
    if (host_fill[k]) memset(host_fill[k], 0, sizeof(double) * (size_t)P * Fo * A);
  MWRT_CUDA_OK(cudaStreamSynchronize(pl->s_d2h));
  for (int k = 0; k < nstr; ++k) MWRT_CUDA_OK(cudaStreamSynchronize(pl->s_cmp[k]));
  if (host_fill[0] || host_fill[1]) {
    const size_t row = (size_t)Fo * A;
    const double qnan = nan("");
    for (int p = 0; p < P; ++p)
      if (st_host[p] & (MWRT_ST_NAN_INPUT | MWRT_ST_RANGE | MWRT_ST_NEG_ABS))   // kernel 3's `bad`
        for (int k = 0; k < 2; ++k)
          if (host_fill[k])
            for (size_t i = 0; i < row; ++i) host_fill[k][(size_t)p * row + i] = qnan;
  }
  if (trace) {
    fprintf(stderr, "mwrt trace: all work queued after %.3f ms of host time\n",
            (ts1.tv_sec - ts0.tv_sec) * 1e3 + (ts1.tv_nsec - ts0.tv_nsec) * 1e-6);
    fprintf(stderr, "mwrt trace (ms since first H2D was queued): chunk size | h2d abs rte d2h done at\n");
    for (int c = 0; c < nchunks; ++c) {
      float t[4] = {0, 0, 0, 0};
      cudaEventElapsedTime(&t[0], tev0, tev[c]);
      for (int k = 1; k < 4; ++k) cudaEventElapsedTime(&t[k], tev0, tev[nchunks + 3 * c + k - 1]);
      fprintf(stderr, "  %d %6d | %.3f %.3f %.3f %.3f\n", c, sizes[c], t[0], t[1], t[2], t[3]);
    }
    for (cudaEvent_t e : tev) cudaEventDestroy(e);
    cudaEventDestroy(tev0);
  }
  return 0;
}

int mwrt_stats_f64(int T, int64_t M, const double* values, const double* refs, double* bias,
                   double* std, double* rmse, double* n_valid, void* stream) {
  if (T <= 0 || M <= 0 || !values || !refs || !bias || !std || !rmse)
    return fail_arg(MWRT_E_ARG, "mwrt_stats_f64: bad argument");
  return stats_launch(T, M, values, refs, bias, std, rmse, n_valid, (cudaStream_t)stream);
}

int mwrt_thin_mean_f64(int S, int N, int M, const double* h, const double* t, const double* p,
                       const double* rh, const int32_t* inds, const int32_t* n_inds,
                       const double* z_offset, double* z_out, double* t_out, double* p_out,
                       double* rh_out, void* stream) {
  if (S <= 0 || N <= 0 || M <= 0 || !h || !t || !p || !rh || !inds || !n_inds || !z_out || !t_out ||
      !p_out || !rh_out)
    return fail_arg(MWRT_E_ARG, "mwrt_thin_mean_f64: bad argument");
  return thin_mean_launch(S, N, M, h, t, p, rh, inds, n_inds, z_offset, z_out, t_out, p_out, rh_out,
                          (cudaStream_t)stream);
}

int mwrt_assemble_levels_f64(int S, int M, int K, int n_levels, double min_p, const double* z,
                             const double* t, const double* p, const double* rh, const int32_t* n_inds,
                             const double* cz_km, const double* cp, const double* ct, const double* cgkg,
                             const double* crh, double* o_p, double* o_t, double* o_ppmv, double* o_h2o,
                             double* o_z, double* o_rh, int32_t* ok, void* stream) {
  if (S <= 0 || M <= 0 || K <= 0 || n_levels < 180 || !z || !t || !p || !rh || !n_inds || !cz_km || !cp ||
      !ct || !cgkg || !crh || !o_p || !o_t || !o_ppmv || !o_h2o || !o_z || !o_rh)
    return fail_arg(MWRT_E_ARG, "mwrt_assemble_levels_f64: bad argument (n_levels >= 180 required)");
  return assemble_launch(S, M, K, n_levels, min_p, z, t, p, rh, n_inds, cz_km, cp, ct, cgkg, crh, o_p,
                         o_t, o_ppmv, o_h2o, o_z, o_rh, ok, (cudaStream_t)stream);
}

int mwrt_cloud_features_f64(int S, int L, const double* p, const double* t, const double* z,
                            const double* rh, double* lwc_kg_m3, double* lwc_kg_kg, double* iwc_kg_m3,
                            double* iwc_kg_kg, double* lwp, double* iwp, void* stream) {
  if (S <= 0 || L <= 0 || !p || !t || !z || !rh || !lwc_kg_m3 || !lwc_kg_kg || !iwc_kg_m3 || !iwc_kg_kg)
    return fail_arg(MWRT_E_ARG, "mwrt_cloud_features_f64: bad argument");
  return cloud_features_launch(S, L, p, t, z, rh, lwc_kg_m3, lwc_kg_kg, iwc_kg_m3, iwc_kg_kg, lwp, iwp,
                               (cudaStream_t)stream);
}

int mwrt_math_selftest_f64(int op, int64_t n, const double* x, double* y, void* stream) {
  if (n <= 0 || !x || !y || op < 0 || op > MWRT_MATH_ATAN2_M1X)
    return fail_arg(MWRT_E_ARG, "mwrt_math_selftest_f64: bad argument");
  return math_selftest_launch(op, n, x, y, (cudaStream_t)stream);
}

int mwrt_measure_dfma_peak(double* flops_per_s, double* sm_clock_mhz_est) {
  if (!flops_per_s) return fail_arg(MWRT_E_ARG, "mwrt_measure_dfma_peak: null argument");
  return dfma_peak(flops_per_s, sm_clock_mhz_est);
}

}  // extern "C"
