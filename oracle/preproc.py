"""CPU restatement of the numerical core of the reference's profile pre-processing.
TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows /root/reference/python_src/preproc/preprocessing4all.py:
  running_mean_from_arrays :156-169     thinning pattern :282-285      humidity helpers :104-130
  jump check :315-331                   add_clim2profiles :478-531     lowest2tenhPa :450-474
  lowest-180 selection :807-814
File readers, cloud features and the MWR matching around these steps are out of scope.  The
climatology is an INPUT here (the reference takes it from pyrtlib's gl_atm(atm=1), which is absent).
"""
import numpy as np

MIN_P = 137.0          # preprocessing4all.py:43
DATAPOINTS_BL = 80     # :44
DATAPOINTS_FT = 120    # :45
N_LEVELS = 180         # :42


def clausius_clapeyron_liq(temp_celsius):
    return 610.78 * np.exp(2.5e6 / 462 * (1 / 273.15 - 1 / (273.15 + temp_celsius)))


def rh2mixing_ratio(RH, abs_T, p):
    e = clausius_clapeyron_liq(abs_T - 273.15) * RH / 100
    q = (0.622 * e) / (p - 0.3777 * e)
    return q / (1 - q)


def rh2ppmv(RH, abs_T, p):
    e = clausius_clapeyron_liq(abs_T - 273.15) * RH / 100
    return 1000000 * e / p


def detect_crop(height):
    """Launch-hold crop of :229-241: count leading samples whose height changes by < 2 m."""
    crop = 0
    old_h = height[0]
    for i in range(1000):
        current_h = height[i]
        if abs(current_h - old_h) < 2.0:
            if i != 0:
                crop += 1
        else:
            break
        old_h = current_h
    return crop


def thinning_indices(height, pressure_hpa, crop=None, min_p=MIN_P, datapoints_bl=DATAPOINTS_BL,
                     datapoints_ft=DATAPOINTS_FT):
    """Index pattern of :250-285: ~80 points below 3000 m, ~120 above, up to the p >= min_p top.
    crop=None: the reference's crop==0 branch (detect the launch hold).  Returns None where the
    reference rejects the sonde (:246-255: max_index < 300 or 10 km not reached)."""
    max_index = int(np.nanargmax(height))
    if pressure_hpa[max_index] < min_p:
        max_index = int(np.nanargmin(np.abs(pressure_hpa[:max_index] - min_p)))
    index3000 = int(np.nanargmin(np.abs(height[:max_index] - 3000)))
    if crop is None:
        crop = detect_crop(height)
    if max_index < 300 or np.nanmax(height) < 10000:
        return None
    increment_bl = int(np.ceil((index3000 - crop) / datapoints_bl))
    increment_ft = int(np.ceil((max_index - index3000) / datapoints_ft))
    inds = np.r_[crop:index3000:increment_bl, index3000:max_index:increment_ft]
    return np.unique(inds)


def running_mean_from_arrays(inds, any_array):
    new_array = []
    ind_max = len(inds)
    for i, ind in enumerate(inds):
        if i == 0 or i == ind_max - 1:
            new_array.append(any_array[inds[i]])
        else:
            lower = int((ind + inds[i - 1]) / 2)
            upper = int((ind + inds[i + 1]) / 2)
            with np.errstate(all="ignore"):
                seg = any_array[lower:upper]
                new_array.append(np.nanmean(seg) if np.isfinite(seg).any() else np.nan)
    return np.array(new_array)


def lowest2tenhPa(p, t, ppmv, m, z, rh):
    if p[179] > 10:
        n_levels = p.size
        idx_10 = np.where(p < 10)[0][0]
        span = np.arange(173, 180)
        target = np.clip(np.linspace(173, idx_10, span.size).round().astype(int), 0, n_levels - 1)
        for arr in (p, t, ppmv, m, z, rh):
            arr[span] = arr[target]
    return p, t, ppmv, m, z, rh


def add_clim2profiles(p, t, ppmv, m, z, rh, clim, min_p=MIN_P):
    """clim: dict z_km, p, t, gkg, rh (50 levels, ground-first).  Returns arrays TOP-first."""
    p_index = np.nanargmin(p)
    wv_index = np.where(ppmv <= 2 * np.nanmin(ppmv))[0][-1]
    z_index = np.nanargmax(z)
    p_threshold = p[int(np.nanmin(np.array([p_index, wv_index, z_index])))]
    p_threshold = min(max(p_threshold, min_p), 200.0)
    mask_clim = p_threshold > clim["p"]
    mask_rs = p > p_threshold
    p = np.concatenate([p[mask_rs], clim["p"][mask_clim]])
    t = np.concatenate([t[mask_rs], clim["t"][mask_clim]])
    m = np.concatenate([m[mask_rs], clim["gkg"][mask_clim] / 1000])
    z = np.concatenate([z[mask_rs], (clim["z_km"] * 1000)[mask_clim]])
    rh = np.concatenate([rh[mask_rs], clim["rh"][mask_clim]])
    ppmv = np.array([rh2ppmv(a, b, c * 100) for a, b, c in zip(rh, t, p)])
    p, t, ppmv, m, z, rh = lowest2tenhPa(p, t, ppmv, m, z, rh)
    return p[::-1], t[::-1], ppmv[::-1], m[::-1], z[::-1], rh[::-1]


def assemble_levels(height, temp, pres_hpa, hum, inds, clim, z_offset=0.0, n_levels=N_LEVELS):
    """One sonde: raw arrays + thinning indices -> the six Level_* columns (n_levels,), top-first
    (NaN-filled when the reference rejects the profile for a z/p jump)."""
    z = height[inds] - z_offset
    t = running_mean_from_arrays(inds, temp)
    p = running_mean_from_arrays(inds, pres_hpa)
    rh = running_mean_from_arrays(inds, hum)
    if np.all(np.array(rh <= 1.5)):
        rh = rh * 100
    m = np.array([rh2mixing_ratio(a, b, c * 100) for a, b, c in zip(rh, t, p)])
    ppmv = np.array([rh2ppmv(a, b, c * 100) for a, b, c in zip(rh, t, p)])
    nan = np.full(n_levels, np.nan)
    for i in range(1, int(len(z) / 2)):
        if abs(z[i] - z[i - 1]) > 500 or abs(p[i] - p[i - 1]) > 50:
            return dict(p=nan, t=nan, ppmv=nan, h2o_gkg=nan, z=nan, rh=nan, ok=False)
    p, t, ppmv, m, z, rh = add_clim2profiles(p, t, ppmv, m, z, rh, clim)

    def low(a):
        a = a[-n_levels:]
        return np.concatenate([np.full(n_levels - len(a), np.nan), a])
    return dict(p=low(p), t=low(t), ppmv=low(ppmv), h2o_gkg=low(1000 * m), z=low(z), rh=low(rh), ok=True)
