/* mwrt.h -- C ABI of libmwrt.so: the B200 (sm_100a) line-by-line microwave forward model.
 *
 * This is the drop-in boundary for the one hot path of the reference: the PyRTlib
 * `TbCloudRTE` simulation driven by
 *     /root/reference/python_src/proc/PyRTlib_processing.py:83-197   (derive_TBs4PyRTlib)
 *     /root/reference/python_src/proc/old_processing.py:520-578
 *     /root/reference/python_src/old_run_pyrtlib/run_pyrtlib_on_all.py:45-58
 * The reference has no FFI: it binds a Python class of a third-party package
 * (`from pyrtlib.tb_spectrum import TbCloudRTE`, PyRTlib_processing.py:28).  The entry points below are
 * what a ctypes binding of that class needs; INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - plain C types only; every array is contiguous fp64, row-major; profiles are GROUND-FIRST
 *     (index 0 = surface), the orientation the reference passes (`z_in[::-1]`, PyRTlib_processing.py:123)
 *   - units as in PyRTlib: z [km], p [mb = hPa], t [K], rh [fraction 0..1], frq [GHz],
 *     angles [deg elevation, 90 = zenith], liquid/ice density [g m^-3], absorption [Np km^-1]
 *   - return value: 0 = ok, <0 = invalid argument (MWRT_E_*), >0 = cudaError_t; text via mwrt_last_error()
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); no entry point
 *     synchronises the device unless its comment says so
 *   - inputs are never modified (the reference reuses its arrays for four model calls in a row,
 *     PyRTlib_processing.py:123,131,139,147)
 *   - a profile with a NaN in z/p/t/rh yields NaN in every output of that profile and nothing else
 *     (the reference skips such profiles and leaves NaN: PyRTlib_processing.py:71-79,117-119,153-154)
 */
#ifndef MWRT_H_
#define MWRT_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MWRT_VERSION 200 /* 0.2.0: mwrt_profiles / mwrt_outputs grew (cloud levels, tmrcld); fp32 line-shape mode */

/* error codes (negative) */
#define MWRT_E_ARG (-1)      /* null pointer / non-positive size / bad flag */
#define MWRT_E_MODEL (-2)    /* unknown model string (PyRTlib raises ValueError in init_absmdl) */
#define MWRT_E_NOMEM (-3)
#define MWRT_E_STATE (-4)    /* plan used with a shape it was not created for */

/* per-profile status bits written by the kernels (int32 per profile) */
#define MWRT_ST_NAN_INPUT 1  /* NaN in z/p/t/rh -> outputs NaN (reference gate) */
#define MWRT_ST_NEG_ABS 2    /* negative absorption met in the layer integration -> outputs NaN
                                (PyRTlib's exponential_integration raises ValueError) */
#define MWRT_ST_DUCTING 4    /* ray tracing hit a non-physical argument; remaining ds = 0 (PyRTlib warns) */
#define MWRT_ST_RANGE 8      /* temperature outside (50 K, 1000 K) -> outputs NaN (the reference's realism
                                window is 0..400 K, preprocessing4all.py:553-572) */

/* flags of mwrt_plan_create */
#define MWRT_F_RAY_TRACING 1 /* refractive slant path (TbCloudRTE(ray_tracing=True)); default 1/sin(elev) */
#define MWRT_F_FROM_SAT 2    /* upwelling (TbCloudRTE.satellite=True); the reference sets False (:125) */
#define MWRT_F_CLOUDY 4      /* liquid/ice absorption from denliq/denice (TbCloudRTE.init_cloudy) */
#define MWRT_F_TAU_LAYERS 8  /* also return the per-layer optical depth (P,F,L) along angle index 0 */

/* column order of the line tables inside mwrt_model_desc (each column = n_lines doubles) */
enum { MWRT_H2O_FL = 0, MWRT_H2O_S1, MWRT_H2O_B2, MWRT_H2O_W0, MWRT_H2O_X, MWRT_H2O_W0S, MWRT_H2O_XS,
       MWRT_H2O_SR, MWRT_H2O_SH, MWRT_H2O_XH, MWRT_H2O_SHS, MWRT_H2O_XHS, MWRT_H2O_AAIR,
       MWRT_H2O_ASELF, MWRT_H2O_W2, MWRT_H2O_W2S, MWRT_H2O_D2, MWRT_H2O_D2S, MWRT_H2O_NCOL };
enum { MWRT_O2_F = 0, MWRT_O2_S300, MWRT_O2_BE, MWRT_O2_W300, MWRT_O2_Y0, MWRT_O2_Y1, MWRT_O2_G0,
       MWRT_O2_G1, MWRT_O2_DNU0, MWRT_O2_DNU1, MWRT_O2_NCOL };

#define MWRT_MAX_H2O_LINES 32
#define MWRT_MAX_O2_LINES 64

/* One absorption model = what `TbCloudRTE.init_absmdl(mdl)` selects (PyRTlib_processing.py:124).
 * Never branched on by name inside a kernel; see tools/gen_model_tables.py for the provenance of the
 * built-in tables (UNVERIFIED recollection of the PyRTlib 1.1.x line lists). */
typedef struct mwrt_model_desc {
  int32_t n_h2o, n_o2;
  int32_t h2o_sd;      /* speed-dependent shape for H2O lines with w2 > 0 */
  int32_t o2_mixing;   /* 1 = first-order (R98..R17), 2 = second-order (R19+) line mixing */
  int32_t o2_dens118;  /* 1998-era separate density for the 118.75 GHz line width */
  int32_t o2_sd118;    /* speed-dependent 118.75 GHz line */
  int32_t n2_fdep;     /* frequency-dependence factor of the N2 continuum */
  int32_t liq_model;   /* 0 = Liebe-91 double Debye, 1 = Rosenkranz-2015 */
  double h2o_reftline, h2o_reftcon, h2o_cf, h2o_xcf, h2o_cs, h2o_xcs, h2o_pvap_div, h2o_den_const,
      h2o_line_const;
  double o2_x, o2_wb300, o2_pvap_div, o2_h2o_broad, o2_nonres, o2_abs_num, o2_abs_den, o2_iso;
  double n2_l, n2_m, n2_n;
  const double* h2o_lines; /* host pointer: MWRT_H2O_NCOL columns x n_h2o */
  const double* o2_lines;  /* host pointer: MWRT_O2_NCOL columns x n_o2 */
} mwrt_model_desc;

typedef struct mwrt_model mwrt_model; /* immutable, device-resident; safe to share between threads */
typedef struct mwrt_plan mwrt_plan;   /* per-caller workspace + streams; one thread at a time */

int mwrt_version(void);
const char* mwrt_last_error(void); /* thread-local text of the last failure */
/* libmwrt links the CUDA runtime statically, so its "current device" is its own: call this before
 * creating models/plans or launching on a device other than 0 (one process per GPU: LOCAL_RANK). */
int mwrt_set_device(int device);
int mwrt_get_device(int* device);

/* ---- models: replaces TbCloudRTE.init_absmdl(str) (PyRTlib_processing.py:124,132,140,148) ---- */
int mwrt_builtin_model_count(void);
const char* mwrt_builtin_model_name(int i);
/* fills *out with pointers into the library's constant tables (valid for the process lifetime) */
int mwrt_builtin_model_desc(const char* name, mwrt_model_desc* out);
int mwrt_model_create(const char* name, mwrt_model** out);                  /* MWRT_E_MODEL if unknown */
int mwrt_model_create_custom(const mwrt_model_desc* desc, mwrt_model** out); /* e.g. true PyRTlib tables */
int mwrt_model_destroy(mwrt_model* m);

/* ---- stage kernels (all pointers DEVICE memory) ------------------------------------------------
 * Exposed so that each stage can be checked against the oracle at the tolerance BASELINE.json states
 * (absorption and optical depth 1e-9 relative, TB 1e-3 K). */

/* RTEquation.vapor: e = rh*es(T) [mb] (Goff-Gratch over water), rho [g m^-3]; n = P*L elements.
 * rho may be NULL. */
int mwrt_vapor_f64(int64_t n, const double* tk, const double* rh, double* e, double* rho, void* stream);

/* RTEquation.clearsky_absorption (+ cloudy_absorption when denliq/denice are given) for every
 * profile x frequency x level.  p,tk,e: (P,L).  denliq,denice: (P,L) or NULL.  frq: (F,) device.
 * awet,adry: (P,F,L) out.  aliq,aice: (P,F,L) out or NULL.  status: (P,) int32, OR-ed, or NULL. */
int mwrt_absorption_f64(const mwrt_model* m, int P, int L, int F, const double* p, const double* tk,
                        const double* e, const double* denliq, const double* denice, const double* frq,
                        double* awet, double* adry, double* aliq, double* aice, int32_t* status,
                        void* stream);
/* Same call in the fp32 LINE-SHAPE mode (BASELINE.json north_star: "any fp32 line-shape mode must stay
 * within 0.01 K and is reported separately"): the per-(line, frequency) Van Vleck-Weisskopf sums run on
 * the FP32 pipe; line parameters per (level, line), continua, the speed-dependent / narrow-line
 * corrections and every output stay fp64.  Not PyRTlib-exact: relative error of awet/adry ~1e-6. */
int mwrt_absorption_f32ls(const mwrt_model* m, int P, int L, int F, const double* p, const double* tk,
                          const double* e, const double* denliq, const double* denice, const double* frq,
                          double* awet, double* adry, double* aliq, double* aice, int32_t* status,
                          void* stream);

/* Slant path length per layer, ds (P,A,L), ds[...,0] = 0.  ray_tracing = 0: diff(z)/sin(elev)
 * (what the reference runs: the 6-argument constructor keeps ray_tracing=False); 1: RTEquation.refractivity
 * + RTEquation.ray_tracing.  z is absolute height; the antenna height z[:,0] is handled inside.
 * angles: (A,) device.  p,tk,e are only read when ray_tracing = 1. */
int mwrt_slant_path_f64(int P, int L, int A, const double* z, const double* p, const double* tk,
                        const double* e, const double* angles, int ray_tracing, double* ds,
                        int32_t* status, void* stream);

/* exponential_integration + planck + bright for every profile x frequency x angle.
 * tk (P,L); awet,adry (P,F,L); aliq,aice (P,F,L) or NULL; ds (P,A,L); frq (F,) device.
 * Outputs (P,F,A), any may be NULL: tbtotal,tbatm,tmr [K]; tauwet,taudry,tauliq,tauice [Np].
 * tau_layer: (P,F,L) total (wet+dry+liquid+ice) optical depth of every layer along the path of
 * angle index `tau_layer_angle` (0 <= index < A), or NULL. */
int mwrt_rte_f64(int P, int L, int F, int A, const double* tk, const double* awet, const double* adry,
                 const double* aliq, const double* aice, const double* ds, const double* frq,
                 int from_sat, double emissivity, double* tbtotal, double* tbatm, double* tmr,
                 double* tauwet, double* taudry, double* tauliq, double* tauice, double* tau_layer,
                 int tau_layer_angle, int32_t* status, void* stream);

/* Channel band-pass average: out[p,c,a] = sum_s w[c,s] * in[p, c*S+s, a]; w (C,S) device, rows sum to 1
 * (BASELINE.json config 3; the reference itself is monochromatic, PyRTlib_processing.py:87-88). */
int mwrt_bandpass_f64(int P, int C, int S, int A, const double* in, const double* w, double* out,
                      void* stream);

/* ---- the fused path: replaces TbCloudRTE(...).execute() for a batch of profiles ----------------
 * A plan fixes (model, L, F, A, frq, angles, flags) and owns device copies of frq/angles, the
 * intermediate (P_chunk,F,L) absorption workspace, and -- for the host entry point -- streams, events
 * and staging buffers.  frq/angles are HOST pointers here.  chunk_profiles <= 0 picks a default.
 * Threading: mwrt_model handles are immutable and may be shared; a plan owns its workspace, so ONE
 * forward call at a time per plan (use one plan per host thread / stream). */
int mwrt_plan_create(const mwrt_model* m, int L, int F, int A, const double* frq, const double* angles,
                     int flags, int chunk_profiles, mwrt_plan** out);
int mwrt_plan_destroy(mwrt_plan* plan);
int mwrt_plan_set_emissivity(mwrt_plan* plan, double es);
/* arithmetic of kernel 1's line sums: MWRT_PREC_F64 (default; the parity path) or
 * MWRT_PREC_F32_LINESHAPE (see mwrt_absorption_f32ls; TB within 0.01 K of the fp64 path) */
#define MWRT_PREC_F64 0
#define MWRT_PREC_F32_LINESHAPE 1
int mwrt_plan_set_precision(mwrt_plan* plan, int precision);
/* Kernel 1 alone on a plan's frequency list, model and precision: awet, adry (P,F,L) from p, tk, e (P,L),
 * all DEVICE memory, status (P,) OR-ed or NULL.  Same results as mwrt_absorption_f64 / _f32ls, but the
 * per-frequency-block tables kernel 1 reads are the plan's cached ones instead of being rebuilt per call
 * (the stage of TbCloudRTE.execute() that PyRTlib runs as RTEquation.clearsky_absorption per angle and
 * frequency; reached by the reference at python_src/proc/PyRTlib_processing.py:126).  This is the call
 * bench.py times for the roofline figure. */
int mwrt_plan_absorption(mwrt_plan* plan, int P, const double* p, const double* tk, const double* e,
                         double* awet, double* adry, int32_t* status, void* stream);
/* pieces > 1: mwrt_plan_forward_device cuts the batch into `pieces` parts and overlaps kernel 3 of
 * part c (on a stream owned by the plan) with kernel 1 of part c+1 (on the caller's stream); the
 * caller's stream still completes after everything.  The host entry point runs equal chunks of ~1000
 * profiles (2*pieces chunks once pieces > 1) on two alternating compute streams; MWRT_TRACE=1 in the
 * environment prints its per-chunk timeline.  Default 1 (no split). */
int mwrt_plan_set_split(mwrt_plan* plan, int pieces);
/* Multi-GPU gather fused into the result stores ("RTE epilogue writes through peer
 * pointers").  The (P,F,A) result pointers handed to mwrt_plan_forward_device are then written
 *   mode MWRT_OUT_PLAIN      as ordinary device memory (default);
 *   mode MWRT_OUT_MULTICAST  with NVLink multicast stores: the pointers are MULTICAST addresses of a
 *                            symmetric buffer, NVSwitch replicates each store to every rank;
 *   mode MWRT_OUT_PEERS      once per peer: plain stores to pointer + peer_byte_offsets[r], r < npeer
 *                            (<= MWRT_MAX_PEERS; include 0 for the own copy).
 * status and tau_layer are always written plainly.  The caller orders the ranks (a barrier after the
 * step) before anyone reads the gathered buffers. */
#define MWRT_OUT_PLAIN 0
#define MWRT_OUT_MULTICAST 1
#define MWRT_OUT_PEERS 2
#define MWRT_MAX_PEERS 16
int mwrt_plan_set_output_fanout(mwrt_plan* plan, int mode, int npeer, const int64_t* peer_byte_offsets);
/* Multi-GPU gather on the COPY ENGINES (the default of bench.py --gpus N): one asynchronous
 * device-to-device copy of `bytes` from `src` (this GPU) to `dst` -- a peer GPU's buffer mapped into this
 * process (symmetric memory / cudaIpc / cuMemMap) -- on `stream`.  A rank pushes its finished (P,F,A)
 * fields to each peer with one such call per peer on a side stream while the next batch computes; no SM
 * runs a collective kernel.  (One plain cudaMemcpyAsync per copy: the batched-memcpy entry points of
 * CUDA 12.8+ are deliberately not used.) */
int mwrt_push_async(void* dst, const void* src, int64_t bytes, void* stream);
/* Multi-GPU gather as a small PUSH KERNEL over NVLink (the default of bench.py --gpus N): `nblocks` CTAs
 * (<= 0: 8) copy `bytes` (multiple of 16, 16-byte aligned) from `src` on this GPU
 *   - to `mc_dst`, the NVLink MULTICAST address of the same offset of a symmetric buffer: one
 *     `multimem.st` per 16 bytes, NVSwitch replicates it into every rank's copy (a rank sends ONCE); or,
 *   - when mc_dst is NULL, to each of the `npeer` peer-mapped addresses in peer_dst (one store per peer).
 * Asynchronous on `stream`; the caller orders the ranks afterwards (symmetric-memory barrier). */
int mwrt_push_kernel(const void* src, int64_t bytes, void* mc_dst, const int64_t* peer_dst, int npeer,
                     int nblocks, void* stream);
/* optional band-pass: C channels x S sub-frequencies (C*S == F), host weights (C,S) or NULL = 1/S.
 * Outputs of the forward calls become (P,C,A).  The average is taken INSIDE kernel 3 (fixed summation
 * order, bit-identical to mwrt_bandpass_f64 on the monochromatic fields) whenever a channel's S
 * sub-frequencies fit one CTA (S <= 384), the path is plane-parallel and the stores are plain;
 * otherwise the plan falls back to monochromatic fields + mwrt_bandpass_f64. */
int mwrt_plan_set_bandpass(mwrt_plan* plan, int C, int S, const double* weights);
/* number of kernels the plan launches for P profiles (bench.py reports it as gpu_launches) */
int mwrt_plan_launch_count(const mwrt_plan* plan, int P);

typedef struct mwrt_profiles { /* (P,L) each; denliq/denice may be NULL unless MWRT_F_CLOUDY */
  const double *z, *p, *tk, *rh, *denliq, *denice;
  /* (P,) level indices of the base and top of the first cloud layer (what TbCloudRTE.init_cloudy's
   * `cldh` selects), -1 = none; only read when mwrt_outputs.tmrcld is asked for; may be NULL */
  const int32_t *cld_ibase, *cld_itop;
} mwrt_profiles;
typedef struct mwrt_outputs { /* (P,F,A) each [(P,C,A) with band-pass]; any may be NULL */
  double *tbtotal, *tbatm, *tmr, *tauwet, *taudry, *tauliq, *tauice;
  double* tau_layer; /* (P,F,L) total layer optical depth at angle 0, needs MWRT_F_TAU_LAYERS */
  int32_t* status;   /* (P,) */
  /* (P,F,A) mean radiating temperature of the first cloud layer (column `tmrcld` of upstream's
   * DataFrame; 0 where there is no cloud).  Device entry point, downwelling, no band-pass. */
  double* tmrcld;
} mwrt_outputs;

/* device pointers in, device pointers out; asynchronous on `stream` */
int mwrt_plan_forward_device(mwrt_plan* plan, int P, const mwrt_profiles* in, const mwrt_outputs* out,
                             void* stream);
/* HOST pointers in, HOST pointers out (pinned memory makes the copies asynchronous): chunks the
 * batch, overlaps H2D / kernels / D2H on the plan's streams, and returns after everything landed. */
int mwrt_plan_forward_host(mwrt_plan* plan, int P, const mwrt_profiles* in, const mwrt_outputs* out);

/* ---- next to the path: model-minus-reference statistics -------------------
 * stats_by_channel (/root/reference/python_src/plot_scripts/multi_campaign_plots_and_ana.py:155-178)
 * for M independent columns (channel x elevation x ...): values, refs (T,M) device; outputs (M,) device.
 * n_valid counts the non-NaN MODEL values (as the reference does); NaN deviations are skipped. */
int mwrt_stats_f64(int T, int64_t M, const double* values, const double* refs, double* bias,
                   double* std, double* rmse, double* n_valid, void* stream);

/* ---- next to the path: profile pre-processing core, batched over S sondes ----
 * /root/reference/python_src/preproc/preprocessing4all.py.  All pointers device memory.
 * mwrt_thin_mean_f64: running_mean_from_arrays (:156-169) over the thinning pattern (:282-285).
 *   raw h,t,p,rh: (S,N) sonde samples; inds: (S,M) int32 thinning indices, n_inds: (S,) valid counts
 *   (M <= 256); z_offset: (S,) metres subtracted from the thinned heights, or NULL.
 *   Out (S,M): thinned z, and segment nan-means of t, p, rh (first/last point: the sample itself). */
int mwrt_thin_mean_f64(int S, int N, int M, const double* h, const double* t, const double* p,
                       const double* rh, const int32_t* inds, const int32_t* n_inds,
                       const double* z_offset, double* z_out, double* t_out, double* p_out,
                       double* rh_out, void* stream);
/* mwrt_assemble_levels_f64: RH scaling and humidity conversions (:104-130, :302-312), z/p jump
 *   rejection (:315-331), climatology top-up above the sonde's pressure threshold (add_clim2profiles
 *   :478-531; climatology levels cz_km,cp,ct,cgkg,crh: (K,) ground-first, K <= 64), lowest2tenhPa
 *   (:450-474), flip to top-first and selection of the lowest n_levels (:807-814).
 *   Out (S,n_levels) top-first: p [hPa], t [K], ppmv, h2o [g/kg], z [m], rh [%]; ok (S,) 0 = rejected. */
int mwrt_assemble_levels_f64(int S, int M, int K, int n_levels, double min_p, const double* z,
                             const double* t, const double* p, const double* rh, const int32_t* n_inds,
                             const double* cz_km, const double* cp, const double* ct, const double* cgkg,
                             const double* crh, double* o_p, double* o_t, double* o_ppmv, double* o_h2o,
                             double* o_z, double* o_rh, int32_t* ok, void* stream);

/* ---- next to the path: batched cloud detection + adiabatic water content --------
 * /root/reference/python_src/preproc/derive_cloud_water.py: derive_cloud_features (:146-363) with
 * calc_lwc (:68-142), for S profiles of L <= 256 levels, TOP-first (index 0 = top) like `Level_*`.
 * p [hPa], t [K], z [m], rh [%] (over liquid): (S,L) device.  Out (S,L): liquid / ice water content in
 * kg m^-3 and kg/kg (`Level_Liquid` / `Level_Ice` are the kg/kg ones); lwp, iwp (S,) [kg m^-2] or NULL. */
int mwrt_cloud_features_f64(int S, int L, const double* p, const double* t, const double* z,
                            const double* rh, double* lwc_kg_m3, double* lwc_kg_kg, double* iwc_kg_m3,
                            double* iwc_kg_kg, double* lwp, double* iwp, void* stream);

/* ---- measurement helper: dependent-free DFMA stream, returns achieved fp64 FLOP/s (2 per DFMA).
 * bench.py uses it as the denominator "of self-measured DFMA peak" (MEASURED_PEAKS.json has no fp64
 * figure).  Synchronises. */
int mwrt_measure_dfma_peak(double* flops_per_s, double* sm_clock_mhz_est);

/* ---- test helper: the device math helpers the kernels are built from (table-driven branch-free
 * exp/log, MUFU+Newton reciprocal, branch-free atan2), evaluated elementwise so tests can check them against
 * libm.  x, y: device, n doubles.  Domain as documented in csrc/mwrt_common.cuh. */
#define MWRT_MATH_EXP_TB 1
#define MWRT_MATH_LOG_TB 3
#define MWRT_MATH_RCP 4
#define MWRT_MATH_RCP_SEED 5 /* raw MUFU.RCP64H */
#define MWRT_MATH_RCP_NR1 6  /* seed + one quadratic Newton step */
#define MWRT_MATH_ATAN2_X1 7 /* atan2_nb(x, 1)  */
#define MWRT_MATH_ATAN2_XM1 8 /* atan2_nb(x, -1) */
#define MWRT_MATH_ATAN2_1X 9 /* atan2_nb(1, x)  */
#define MWRT_MATH_ATAN2_M1X 10 /* atan2_nb(-1, x) */
int mwrt_math_selftest_f64(int op, int64_t n, const double* x, double* y, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MWRT_H_ */
