"""N1: profile pre-processing on the device -- the step immediately before the hot
path, which defines its input contract (`Level_*`, 180 levels, top-first).

Replaces, batched over sondes, the numerical core of
`/root/reference/python_src/preproc/preprocessing4all.py`: thinning + running means (:156-169,
:282-301), humidity conversions (:104-130), the z/p jump rejection (:315-331), the climatology top-up
`add_clim2profiles` (:478-531), `lowest2tenhPa` (:450-474) and the flip / lowest-180 selection
(:807-814).  File readers (three sonde formats), cloud features and MWR matching stay out of scope.

Climatology: the reference calls pyrtlib's `gl_atm(atm=1)` (AFGL mid-latitude summer) +
`ppmv2gkg` + `mr2rh` (:503-505).  pyrtlib is absent, so the AFGL tables below are written from memory
(UNVERIFIED digits, same status as SURVEY.md Appendix C) and the climatology is an explicit argument.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .forward import _need_cuda, _ptr, _stream

MIN_P, DATAPOINTS_BL, DATAPOINTS_FT, N_LEVELS = 137.0, 80, 120, 180   # preprocessing4all.py:42-45

_Z_KM = np.array(list(range(0, 26)) + [27.5, 30.0, 32.5, 35.0, 37.5, 40.0, 42.5, 45.0, 47.5, 50.0, 55.0,
                                        60.0, 65.0, 70.0, 75.0, 80.0, 85.0, 90.0, 95.0, 100.0, 105.0, 110.0,
                                        115.0, 120.0], dtype=np.float64)
# AFGL-TR-86-0110 mid-latitude summer (gl_atm(atm=1)) -- FROM MEMORY, verify before scientific use
_MLS_P = np.array([1013., 902., 802., 710., 628., 554., 487., 426., 372., 324., 281., 243., 209., 179.,
                   153., 130., 111., 95., 81.2, 69.5, 59.5, 51., 43.7, 37.6, 32.2, 27.7, 19.07, 13.2,
                   9.3, 6.52, 4.64, 3.33, 2.41, 1.76, 1.29, 0.951, 0.515, 0.272, 0.139, 0.067, 0.03,
                   0.012, 4.48e-3, 1.64e-3, 6.25e-4, 2.58e-4, 1.17e-4, 6.11e-5, 3.56e-5, 2.27e-5])
_MLS_T = np.array([294.2, 289.7, 285.2, 279.2, 273.2, 267.2, 261.2, 254.7, 248.2, 241.7, 235.3, 228.8,
                   222.3, 215.8, 215.7, 215.7, 215.7, 215.7, 216.8, 217.9, 219.2, 220.4, 221.6, 222.8,
                   223.9, 225.1, 228.45, 233.7, 239.0, 245.2, 251.3, 257.5, 263.7, 269.9, 275.2, 275.7,
                   269.3, 257.1, 240.1, 218.1, 196.1, 174.1, 165.1, 165.0, 178.3, 190.5, 222.2, 262.4,
                   316.8, 380.0])
_MLS_H2O = np.array([1.876e4, 1.378e4, 9.68e3, 5.984e3, 3.813e3, 2.225e3, 1.51e3, 1.02e3, 6.464e2,
                     4.129e2, 2.472e2, 9.556e1, 2.944e1, 8.0, 5.0, 3.4, 3.3, 3.2, 3.15, 3.2, 3.3, 3.45,
                     3.6, 3.85, 4.0, 4.2, 4.45, 4.7, 4.85, 4.95, 5.0, 5.1, 5.3, 5.45, 5.5, 5.5, 5.35,
                     5.0, 4.4, 3.7, 2.95, 2.1, 1.33, 0.85, 0.54, 0.4, 0.34, 0.28, 0.24, 0.2])


def _goff_gratch(tk):
    y = 373.16 / tk
    es = (-7.90298 * (y - 1.0) + 5.02808 * np.log10(y)
          - 1.3816e-07 * (10 ** (11.344 * (1.0 - (1.0 / y))) - 1.0)
          + 0.0081328 * (10 ** (-3.49149 * (y - 1.0)) - 1.0) + np.log10(1013.246))
    return 10.0 ** es


def midlatitude_summer():
    """dict z_km, p [hPa], t [K], gkg [g/kg], rh [%] -- the chain gl_atm -> ppmv2gkg -> mr2rh(...)[0]
    of preprocessing4all.py:503-505, ground-first, 50 levels."""
    gkg = _MLS_H2O * (18.01528 / 28.9644) * 1e-3
    wsat = 622.0 * _goff_gratch(_MLS_T) / _MLS_P
    return dict(z_km=_Z_KM.copy(), p=_MLS_P.copy(), t=_MLS_T.copy(), gkg=gkg, rh=100.0 * gkg / wsat)


def detect_crop(height):
    """Launch-hold crop (preprocessing4all.py:229-241): leading samples rising by less than 2 m."""
    dh = np.abs(np.diff(np.asarray(height[:1000], dtype=np.float64)))
    stop = np.nonzero(~(dh < 2.0))[0]
    return int(stop[0]) if stop.size else int(dh.size)


def thinning_indices(height, pressure_hpa, crop=None, min_p=MIN_P, datapoints_bl=DATAPOINTS_BL,
                     datapoints_ft=DATAPOINTS_FT):
    """Host-side index logic of preprocessing4all.py:220-285 (data-dependent, O(N) per sonde).
    crop=None detects the launch hold like the reference's crop==0 branch.  Returns None for sondes
    the reference rejects (:246-255: fewer than 300 usable samples, or 10 km not reached)."""
    max_index = int(np.nanargmax(height))
    if pressure_hpa[max_index] < min_p:
        max_index = int(np.nanargmin(np.abs(pressure_hpa[:max_index] - min_p)))
    index3000 = int(np.nanargmin(np.abs(height[:max_index] - 3000)))
    if crop is None:
        crop = detect_crop(height)
    if max_index < 300 or np.nanmax(height) < 10000:
        return None
    inc_bl = int(np.ceil((index3000 - crop) / datapoints_bl))
    inc_ft = int(np.ceil((max_index - index3000) / datapoints_ft))
    return np.unique(np.r_[crop:index3000:inc_bl, index3000:max_index:inc_ft])


def assemble_profiles(height, temp, pres_hpa, hum, inds_list, clim=None, z_offset=None, n_levels=N_LEVELS,
                      min_p=MIN_P):
    """height/temp/pres_hpa/hum: (S,N) raw sonde samples (NaN padded); inds_list: per-sonde index
    arrays from `thinning_indices`.  Returns dict of (S,n_levels) CUDA tensors, top-first:
    Level_Pressure, Level_Temperature, Level_ppmvs, Level_H2O [g/kg], Level_z [m], Level_RH [%], ok."""
    _need_cuda()
    lib = _lib.load()
    clim = clim or midlatitude_summer()
    S, N = np.asarray(height).shape
    M = max([len(i) for i in inds_list if i is not None] + [2])
    if M > 256:
        raise ValueError("at most 256 thinned points per sonde")
    inds = np.zeros((S, M), dtype=np.int32)
    n_inds = np.zeros((S,), dtype=np.int32)
    for s, idx in enumerate(inds_list):
        if idx is None:        # rejected by the host-side checks -> NaN levels, ok = 0
            continue
        inds[s, :len(idx)] = idx
        n_inds[s] = len(idx)

    def dev(x, dt=torch.float64):
        return torch.from_numpy(np.ascontiguousarray(x)).to("cuda", dt).contiguous()
    h, t, p, r = dev(height), dev(temp), dev(pres_hpa), dev(hum)
    di, dn = dev(inds, torch.int32), dev(n_inds, torch.int32)
    zo = dev(z_offset) if z_offset is not None else None
    zt, tt, pt, rt = (torch.empty((S, M), dtype=torch.float64, device="cuda") for _ in range(4))
    _lib.check(lib.mwrt_thin_mean_f64(S, N, M, _ptr(h), _ptr(t), _ptr(p), _ptr(r), _ptr(di), _ptr(dn),
                                      _ptr(zo), _ptr(zt), _ptr(tt), _ptr(pt), _ptr(rt), _stream()),
               "thin_mean")
    c = {k: dev(clim[k]) for k in ("z_km", "p", "t", "gkg", "rh")}
    K = len(clim["p"])
    out = {k: torch.empty((S, n_levels), dtype=torch.float64, device="cuda")
           for k in ("Level_Pressure", "Level_Temperature", "Level_ppmvs", "Level_H2O", "Level_z", "Level_RH")}
    ok = torch.empty((S,), dtype=torch.int32, device="cuda")
    _lib.check(lib.mwrt_assemble_levels_f64(S, M, K, int(n_levels), float(min_p), _ptr(zt), _ptr(tt),
                                            _ptr(pt), _ptr(rt), _ptr(dn), _ptr(c["z_km"]), _ptr(c["p"]),
                                            _ptr(c["t"]), _ptr(c["gkg"]), _ptr(c["rh"]),
                                            _ptr(out["Level_Pressure"]), _ptr(out["Level_Temperature"]),
                                            _ptr(out["Level_ppmvs"]), _ptr(out["Level_H2O"]),
                                            _ptr(out["Level_z"]), _ptr(out["Level_RH"]), _ptr(ok), _stream()),
               "assemble_levels")
    torch.cuda.current_stream().synchronize()
    out["ok"] = ok
    out["thinned"] = dict(z=zt, t=tt, p=pt, rh=rt, n=dn)
    return out


def preprocess_sondes(height, temp, pres_hpa, hum, vital=None, clim=None, n_levels=N_LEVELS):
    """Raw sondes (S,N), NaN padded -> Level_* (S,n_levels) top-first: thinning pattern on the host,
    running means + climatology top-up + selection on the device.  `vital`: per-sonde bool, the
    reference subtracts 20 m from z for the Vital-I site (preprocessing4all.py:292-294)."""
    height = np.asarray(height, dtype=np.float64)
    inds = [thinning_indices(height[s], np.asarray(pres_hpa[s], dtype=np.float64))
            for s in range(height.shape[0])]
    zoff = None if vital is None else np.where(np.asarray(vital, dtype=bool), 20.0, 0.0)
    return assemble_profiles(height, temp, pres_hpa, hum, inds, clim=clim, z_offset=zoff,
                             n_levels=n_levels)


def derive_cloud_features(p_hpa, t_k, z_m, rh_pct):
    """Batched `derive_cloud_features` of the reference (derive_cloud_water.py:146-363): (S,L) profiles,
    TOP-first like `Level_*`, RH in % over liquid.  Returns dict of CUDA tensors: lwc_kg_m3, lwc_kg_kg
    (`Level_Liquid`), iwc_kg_m3, iwc_kg_kg (`Level_Ice`) of shape (S,L) and lwp, iwp (S,) [kg m^-2]."""
    _need_cuda()
    lib = _lib.load()

    def dev(x):
        if isinstance(x, torch.Tensor):
            return x.to("cuda", torch.float64).contiguous()
        return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).cuda()
    p, t, z, rh = dev(p_hpa), dev(t_k), dev(z_m), dev(rh_pct)
    S, L = p.shape
    out = {k: torch.empty((S, L), dtype=torch.float64, device="cuda")
           for k in ("lwc_kg_m3", "lwc_kg_kg", "iwc_kg_m3", "iwc_kg_kg")}
    out["lwp"] = torch.empty((S,), dtype=torch.float64, device="cuda")
    out["iwp"] = torch.empty((S,), dtype=torch.float64, device="cuda")
    _lib.check(lib.mwrt_cloud_features_f64(S, L, _ptr(p), _ptr(t), _ptr(z), _ptr(rh), _ptr(out["lwc_kg_m3"]),
                                           _ptr(out["lwc_kg_kg"]), _ptr(out["iwc_kg_m3"]),
                                           _ptr(out["iwc_kg_kg"]), _ptr(out["lwp"]), _ptr(out["iwp"]),
                                           _stream()), "cloud_features")
    torch.cuda.current_stream().synchronize()
    return out
