"""Batched forward model: the GPU replacement of `TbCloudRTE(...).execute()` for P profiles at once.

Reference path replaced: the triple loop time x Crop x elevation with four model evaluations each at
`/root/reference/python_src/proc/PyRTlib_processing.py:99-151`.  All arithmetic runs in libmwrt.so
(hand-written sm_100a CUDA behind the C ABI of include/mwrt.h); PyTorch is only used for device
memory, streams and (in `distributed.py`) NCCL.  No CPU fallback.

Conventions (PyRTlib's): profiles ground-first, z [km], p [mb], t [K], rh [fraction], frq [GHz],
angles [deg elevation]; outputs (P, F, A).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .models import get_model

FIELDS = ("tbtotal", "tbatm", "tmr", "tauwet", "taudry", "tauliq", "tauice")


def _ptr(t):
    if t is None:
        return C.c_void_p(None)
    if isinstance(t, int):          # raw device (or NVLink multicast) address
        return C.c_void_p(t)
    return C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("mwr_fast_forward_operators_and_lbls_b200 needs a CUDA device (B200, "
                           "sm_100a); there is no CPU fallback")
    _lib.set_device(torch.cuda.current_device())


def _dev64(x, device=None):
    """(P,L) fp64 contiguous CUDA tensor from tensor / ndarray, never aliasing-then-mutating input."""
    if isinstance(x, torch.Tensor):
        t = x
    else:
        t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64))
    return t.to(device=device or "cuda", dtype=torch.float64).contiguous()


def _host_f64(x):
    return np.ascontiguousarray(np.atleast_1d(np.asarray(x, dtype=np.float64)))


class Plan:
    """`mwrt_plan`: fixes (model, L, F, A, frq, angles, flags); owns workspace and streams."""

    def __init__(self, model, L, frq, angles, ray_tracing=False, from_sat=False, cloudy=False,
                 tau_layers=False, chunk_profiles=0, emissivity=1.0, bandpass=None, f32_lineshape=False):
        _need_cuda()
        self.lib = _lib.load()
        self.model = get_model(model)
        self.frq = _host_f64(frq)
        self.angles = _host_f64(angles)
        if np.isnan(self.frq).any() or np.isnan(self.angles).any():
            raise ValueError("NaN in frequencies or angles")  # reference gate, PyRTlib_processing.py:77-78
        self.L, self.F, self.A = int(L), len(self.frq), len(self.angles)
        self.flags = ((_lib.F_RAY_TRACING if ray_tracing else 0) | (_lib.F_FROM_SAT if from_sat else 0)
                      | (_lib.F_CLOUDY if cloudy else 0) | (_lib.F_TAU_LAYERS if tau_layers else 0))
        self.cloudy, self.tau_layers = bool(cloudy), bool(tau_layers)
        h = C.c_void_p()
        _lib.check(self.lib.mwrt_plan_create(self.model.handle, self.L, self.F, self.A,
                                             self.frq.ctypes.data_as(_lib.c_dp),
                                             self.angles.ctypes.data_as(_lib.c_dp), self.flags,
                                             int(chunk_profiles), C.byref(h)), "plan_create")
        self._h = h
        self.Fo = self.F
        if emissivity != 1.0:
            _lib.check(self.lib.mwrt_plan_set_emissivity(h, float(emissivity)), "set_emissivity")
        if bandpass is not None:
            nchan, nsub, w = bandpass
            wp = None
            if w is not None:
                w = np.ascontiguousarray(w, dtype=np.float64).reshape(nchan, nsub)
                wp = w.ctypes.data_as(_lib.c_dp)
            _lib.check(self.lib.mwrt_plan_set_bandpass(h, int(nchan), int(nsub), wp), "set_bandpass")
            self.Fo = int(nchan)
        self.f32_lineshape = bool(f32_lineshape)
        if f32_lineshape:   # fp32 line-shape mode (north_star: <= 0.01 K, reported separately)
            _lib.check(self.lib.mwrt_plan_set_precision(h, _lib.PREC_F32_LINESHAPE), "set_precision")

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                self.lib.mwrt_plan_destroy(h)
            except Exception:
                pass

    def set_split(self, pieces):
        """Overlap kernel 3 of one part of the batch with kernel 1 of the next (mwrt_plan_set_split)."""
        _lib.check(self.lib.mwrt_plan_set_split(self._h, int(pieces)), "set_split")

    def set_output_fanout(self, mode="plain", peer_byte_offsets=None):
        """How kernel 3 stores the (P,F,A) results (mwrt_plan_set_output_fanout): "plain",
        "multicast" (the `out` entries of `forward` are then NVLink multicast ADDRESSES, given as int)
        or "peers" (one store per byte offset in `peer_byte_offsets`, 0 = own copy)."""
        code = {"plain": 0, "multicast": 1, "peers": 2}[mode]
        offs = list(peer_byte_offsets or [])
        arr = (C.c_int64 * max(len(offs), 1))(*offs)
        _lib.check(self.lib.mwrt_plan_set_output_fanout(self._h, code, len(offs), arr), "set_output_fanout")

    def absorption(self, p, t, e, out=None):
        """Kernel 1 alone on this plan's frequencies/model/precision (`mwrt_plan_absorption`): awet, adry
        (P,F,L) CUDA tensors from p, t, e (P,L) CUDA tensors; the per-block tables are the plan's cached
        ones (the plan-less `absorption()` rebuilds them per call)."""
        P = p.shape[0]
        for x, name in ((p, "p"), (t, "t"), (e, "e")):
            self._check_dev(x, P, name)
        if out is None:
            out = (torch.empty((P, self.F, self.L), dtype=torch.float64, device="cuda"),
                   torch.empty((P, self.F, self.L), dtype=torch.float64, device="cuda"))
        _lib.check(self.lib.mwrt_plan_absorption(self._h, P, _ptr(p), _ptr(t), _ptr(e), _ptr(out[0]),
                                                 _ptr(out[1]), None, _stream()), "plan_absorption")
        return out

    def launch_count(self, P):
        return int(self.lib.mwrt_plan_launch_count(self._h, int(P)))

    # ---- device buffers in, device buffers out ------------------------------------------------
    def _check_dev(self, x, P, what):
        if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float64 and x.is_contiguous()
                and tuple(x.shape) == (P, self.L)):
            raise ValueError("%s must be a contiguous CUDA fp64 tensor of shape (P,%d)" % (what, self.L))

    def forward(self, z, p, t, rh, denliq=None, denice=None, fields=FIELDS, out=None, cld_ibase=None,
                cld_itop=None):
        """All arguments (P,L) CUDA fp64 tensors.  Returns dict of (P,Fo,A) CUDA tensors
        (+ 'tau_layer' (P,F,L) if the plan was made with tau_layers, + 'status' (P,) int32).
        `fields` may name 'tmrcld' (upstream's cloud mean radiating temperature); it needs
        cld_ibase / cld_itop, (P,) int32 CUDA tensors with the level indices of the cloud's base and top
        (-1: none).  Asynchronous on the current torch stream."""
        P = z.shape[0]
        for x, what in ((z, "z"), (p, "p"), (t, "t"), (rh, "rh")):
            self._check_dev(x, P, what)
        if self.cloudy:
            if denliq is None or denice is None:
                raise ValueError("a cloudy plan needs denliq and denice")
            self._check_dev(denliq, P, "denliq")
            self._check_dev(denice, P, "denice")
        dev = z.device
        want_tmrcld = "tmrcld" in fields
        if want_tmrcld:
            for x in (cld_ibase, cld_itop):
                if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.int32
                        and x.is_contiguous() and tuple(x.shape) == (P,)):
                    raise ValueError("tmrcld needs cld_ibase and cld_itop: contiguous CUDA int32 tensors (P,)")
        res = out if out is not None else {}
        for f in fields:
            if f not in res:      # (an int entry is a raw / multicast address, see set_output_fanout)
                res[f] = torch.empty((P, self.Fo, self.A), dtype=torch.float64, device=dev)
            elif isinstance(res[f], torch.Tensor):
                if not (res[f].is_cuda and res[f].dtype == torch.float64 and res[f].is_contiguous()
                        and tuple(res[f].shape) == (P, self.Fo, self.A)):
                    raise ValueError("out['%s'] must be a contiguous CUDA fp64 tensor of shape (%d,%d,%d)"
                                     % (f, P, self.Fo, self.A))
        if self.tau_layers and "tau_layer" not in res:
            res["tau_layer"] = torch.empty((P, self.F, self.L), dtype=torch.float64, device=dev)
        if "status" not in res:
            res["status"] = torch.empty((P,), dtype=torch.int32, device=dev)
        pin = _lib.Profiles(_ptr(z), _ptr(p), _ptr(t), _ptr(rh),
                            _ptr(denliq if self.cloudy else None), _ptr(denice if self.cloudy else None),
                            _ptr(cld_ibase if want_tmrcld else None), _ptr(cld_itop if want_tmrcld else None))
        pout = _lib.Outputs(*[_ptr(res.get(f)) for f in FIELDS],
                            _ptr(res.get("tau_layer")), _ptr(res["status"]),
                            _ptr(res.get("tmrcld") if want_tmrcld else None))
        with torch.cuda.device(dev):
            _lib.set_device(dev.index if dev.index is not None else torch.cuda.current_device())
            _lib.check(self.lib.mwrt_plan_forward_device(self._h, int(P), C.byref(pin), C.byref(pout),
                                                         _stream()), "forward_device")
        return res

    # ---- host buffers in, host buffers out (chunked copy/compute overlap inside the library) ---
    def forward_host(self, z, p, t, rh, denliq=None, denice=None, fields=FIELDS, out=None):
        """All arguments (P,L) host fp64 arrays (numpy or CPU torch; pinned memory makes the copies
        asynchronous).  Returns dict of numpy arrays; blocks until the results are in host memory."""
        def as_np(x):
            if x is None:
                return None
            if isinstance(x, torch.Tensor):
                x = x.numpy()
            return np.ascontiguousarray(x, dtype=np.float64)
        z, p, t, rh, denliq, denice = (as_np(v) for v in (z, p, t, rh, denliq, denice))
        P = z.shape[0]
        for x in (z, p, t, rh):
            if x.shape != (P, self.L):
                raise ValueError("inputs must have shape (P,%d)" % self.L)
        if self.cloudy and (denliq is None or denice is None):
            raise ValueError("a cloudy plan needs denliq and denice")
        for x in (denliq, denice):
            if x is not None and x.shape != (P, self.L):
                raise ValueError("denliq/denice must have shape (P,%d)" % self.L)
        res = out if out is not None else {}
        for f in FIELDS:
            v = res.get(f)
            if v is None:
                continue
            if isinstance(v, torch.Tensor):
                v = v.numpy()
            if not (isinstance(v, np.ndarray) and v.dtype == np.float64 and v.flags["C_CONTIGUOUS"]
                    and v.shape == (P, self.Fo, self.A)):
                raise ValueError("out['%s'] must be a C-contiguous float64 array of shape (%d,%d,%d)"
                                 % (f, P, self.Fo, self.A))
        if "status" in res:
            v = res["status"].numpy() if isinstance(res["status"], torch.Tensor) else res["status"]
            if not (v.dtype == np.int32 and v.flags["C_CONTIGUOUS"] and v.shape == (P,)):
                raise ValueError("out['status'] must be a C-contiguous int32 array of shape (%d,)" % P)
        missing = [f for f in fields if f not in res]
        if missing:   # one pinned block: asynchronous copies, and one 2-D copy per chunk for all fields
            res.update(pinned_fields(missing, (P, self.Fo, self.A)))
        if self.tau_layers and "tau_layer" not in res:
            res["tau_layer"] = np.empty((P, self.F, self.L), dtype=np.float64)
        if "status" not in res:
            res["status"] = np.empty((P,), dtype=np.int32)

        def hp(x):
            if x is None:
                return C.c_void_p(None)
            if isinstance(x, torch.Tensor):
                return C.c_void_p(x.data_ptr())
            return C.c_void_p(x.ctypes.data)
        pin = _lib.Profiles(hp(z), hp(p), hp(t), hp(rh), hp(denliq if self.cloudy else None),
                            hp(denice if self.cloudy else None), hp(None), hp(None))
        pout = _lib.Outputs(*[hp(res.get(f)) for f in FIELDS], hp(res.get("tau_layer")), hp(res["status"]),
                            hp(None))
        _lib.set_device(torch.cuda.current_device())
        _lib.check(self.lib.mwrt_plan_forward_host(self._h, int(P), C.byref(pin), C.byref(pout)),
                   "forward_host")
        return res


def pinned_fields(names, shape, dtype=np.float64):
    """dict name -> numpy view of ONE page-locked block (fields equally spaced): host buffers for
    `Plan.forward_host` that the library moves with a single 2-D copy per chunk.  The block stays
    alive as long as any view does."""
    names = list(names)
    block = torch.empty((len(names),) + tuple(shape), dtype=torch.from_numpy(np.empty(0, dtype)).dtype)
    block = block.pin_memory()
    arr = block.numpy()
    return {n: arr[i] for i, n in enumerate(names)}


def simulate(z, p, t, rh, frq, angles, model, ray_tracing=False, from_sat=False, denliq=None,
             denice=None, fields=FIELDS, tau_layers=False, bandpass=None, device=None,
             f32_lineshape=False):
    """One-call batched forward model (SURVEY.md 8(b) "batched entry").

    z,p,t,rh[,denliq,denice]: (P,L) arrays/tensors, ground-first.  Returns dict of (P,F,A) CUDA
    tensors.  A profile containing NaN gives NaN outputs (status bit 1) and the batch continues --
    the reference's per-profile gate at PyRTlib_processing.py:117-119,153-154."""
    _need_cuda()
    zt = _dev64(z, device)
    cloudy = denliq is not None or denice is not None
    if cloudy:
        denliq = torch.zeros_like(zt) if denliq is None else _dev64(denliq, zt.device)
        denice = torch.zeros_like(zt) if denice is None else _dev64(denice, zt.device)
    plan = Plan(model, zt.shape[1], frq, angles, ray_tracing=ray_tracing, from_sat=from_sat,
                cloudy=cloudy, tau_layers=tau_layers, bandpass=bandpass, f32_lineshape=f32_lineshape)
    res = plan.forward(zt, _dev64(p, zt.device), _dev64(t, zt.device), _dev64(rh, zt.device),
                       denliq, denice, fields=fields)
    torch.cuda.current_stream().synchronize()  # the plan (and its workspace) dies with this frame
    return res


# ------------------------------------------------------------------------------------------------
# stage entry points (used by the parity tests: absorption / optical depth at 1e-9, TB at 1e-3 K)
# ------------------------------------------------------------------------------------------------
def vapor(t, rh):
    """RTEquation.vapor on device: returns e [mb], rho [g m^-3]."""
    _need_cuda()
    lib = _lib.load()
    t, rh = _dev64(t), _dev64(rh)
    e, rho = torch.empty_like(t), torch.empty_like(t)
    _lib.check(lib.mwrt_vapor_f64(t.numel(), _ptr(t), _ptr(rh), _ptr(e), _ptr(rho), _stream()), "vapor")
    return e, rho


def absorption(model, p, t, e, frq, denliq=None, denice=None, f32_lineshape=False):
    """awet, adry[, aliq, aice] as (P,F,L) CUDA tensors + status (P,).  f32_lineshape: the fp32
    line-shape mode (`mwrt_absorption_f32ls`)."""
    _need_cuda()
    lib = _lib.load()
    m = get_model(model)
    p, t, e = _dev64(p), _dev64(t), _dev64(e)
    P, L = p.shape
    f = torch.from_numpy(_host_f64(frq)).cuda()
    F = f.numel()
    awet = torch.empty((P, F, L), dtype=torch.float64, device="cuda")
    adry = torch.empty_like(awet)
    cloudy = denliq is not None or denice is not None
    aliq = aice = None
    if cloudy:
        denliq = torch.zeros_like(p) if denliq is None else _dev64(denliq)
        denice = torch.zeros_like(p) if denice is None else _dev64(denice)
        aliq, aice = torch.empty_like(awet), torch.empty_like(awet)
    status = torch.zeros((P,), dtype=torch.int32, device="cuda")
    fn = lib.mwrt_absorption_f32ls if f32_lineshape else lib.mwrt_absorption_f64
    _lib.check(fn(m.handle, P, L, F, _ptr(p), _ptr(t), _ptr(e), _ptr(denliq), _ptr(denice), _ptr(f),
                  _ptr(awet), _ptr(adry), _ptr(aliq), _ptr(aice), _ptr(status), _stream()), "absorption")
    torch.cuda.current_stream().synchronize()
    return dict(awet=awet, adry=adry, aliq=aliq, aice=aice, status=status)


def slant_path(z, p, t, e, angles, ray_tracing=False):
    """ds (P,A,L) [km] + status."""
    _need_cuda()
    lib = _lib.load()
    z = _dev64(z)
    P, L = z.shape
    a = torch.from_numpy(_host_f64(angles)).cuda()
    A = a.numel()
    p = _dev64(p) if p is not None else None
    t = _dev64(t) if t is not None else None
    e = _dev64(e) if e is not None else None
    ds = torch.empty((P, A, L), dtype=torch.float64, device="cuda")
    status = torch.zeros((P,), dtype=torch.int32, device="cuda")
    _lib.check(lib.mwrt_slant_path_f64(P, L, A, _ptr(z), _ptr(p), _ptr(t), _ptr(e), _ptr(a),
                                       int(bool(ray_tracing)), _ptr(ds), _ptr(status), _stream()),
               "slant_path")
    torch.cuda.current_stream().synchronize()
    return ds, status


def rte(t, awet, adry, ds, frq, aliq=None, aice=None, from_sat=False, emissivity=1.0,
        tau_layer_angle=None):
    """Layer integration + RTE from given absorption and path: dict of (P,F,A) tensors."""
    _need_cuda()
    lib = _lib.load()
    t = _dev64(t)
    P, F, L = awet.shape
    A = ds.shape[1]
    f = torch.from_numpy(_host_f64(frq)).cuda()
    res = {k: torch.empty((P, F, A), dtype=torch.float64, device="cuda") for k in FIELDS}
    tl = None
    if tau_layer_angle is not None:
        tl = torch.empty((P, F, L), dtype=torch.float64, device="cuda")
        res["tau_layer"] = tl
    status = torch.zeros((P,), dtype=torch.int32, device="cuda")
    res["status"] = status
    _lib.check(lib.mwrt_rte_f64(P, L, F, A, _ptr(t), _ptr(awet), _ptr(adry), _ptr(aliq), _ptr(aice),
                                _ptr(ds), _ptr(f), int(bool(from_sat)), float(emissivity),
                                *[_ptr(res[k]) for k in FIELDS], _ptr(tl),
                                int(tau_layer_angle or 0), _ptr(status), _stream()), "rte")
    torch.cuda.current_stream().synchronize()
    return res


def bandpass(x, nchan, nsub, weights=None):
    """(P, nchan*nsub, A) -> (P, nchan, A) weighted channel average."""
    _need_cuda()
    lib = _lib.load()
    x = x.contiguous()
    P, F, A = x.shape
    if F != nchan * nsub:
        raise ValueError("nchan*nsub must equal the frequency axis")
    w = None
    if weights is not None:
        w = torch.from_numpy(np.ascontiguousarray(weights, dtype=np.float64).reshape(nchan, nsub)).cuda()
    out = torch.empty((P, nchan, A), dtype=torch.float64, device="cuda")
    _lib.check(lib.mwrt_bandpass_f64(P, nchan, nsub, A, _ptr(x), _ptr(w), _ptr(out), _stream()),
               "bandpass")
    torch.cuda.current_stream().synchronize()
    return out


def measure_dfma_peak():
    """(fp64 FLOP/s of a DFMA stream, SM clock estimate in MHz) measured on the current device."""
    _need_cuda()
    lib = _lib.load()
    fl, mhz = C.c_double(), C.c_double()
    _lib.check(lib.mwrt_measure_dfma_peak(C.byref(fl), C.byref(mhz)), "measure_dfma_peak")
    return fl.value, mhz.value


MATH_OPS = {"exp_tb": 1, "log_tb": 3, "rcp": 4, "rcp_seed": 5, "rcp_nr1": 6, "atan2_x1": 7, "atan2_xm1": 8,
            "atan2_1x": 9, "atan2_m1x": 10}


def math_selftest(op, x):
    """Evaluate one of the device math helpers (csrc/mwrt_common.cuh) elementwise -- test hook."""
    _need_cuda()
    lib = _lib.load()
    xd = _dev64(x).contiguous()
    y = torch.empty_like(xd)
    _lib.check(lib.mwrt_math_selftest_f64(MATH_OPS[op], xd.numel(), _ptr(xd), _ptr(y), _stream()),
               "math_selftest")
    torch.cuda.current_stream().synchronize()
    return y
