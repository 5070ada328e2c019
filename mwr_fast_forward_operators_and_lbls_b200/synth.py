"""Synthetic radiosonde-shaped profile batches for the BASELINE.json configurations.

The reference's real inputs are 180-level sondes topped up with climatology
(`/root/reference/python_src/preproc/preprocessing4all.py:282-301,478-531`, denser near the ground);
they are not available offline, so benchmarks and parity tests use this generator.  Values respect
the reference's realism windows (`preprocessing4all.py:553-572`: p in [0,1100] hPa, T in [0,400] K,
RH in [0,110] %).  Deterministic for a given (P, L, seed); NumPy on the host so that the oracle and
the GPU consume bit-identical arrays.
"""
import numpy as np

HATPRO_FRQ = np.array([22.24, 23.04, 23.84, 25.44, 26.24, 27.84, 31.4, 51.26, 52.28, 53.86, 54.94,
                       56.66, 57.3, 58.0])                       # PyRTlib_processing.py:87-88
HATPRO_ELEV = np.array([90., 30, 19.2, 14.4, 11.4, 8.4, 6.6, 5.4, 4.8, 4.2])  # PyRTlib_processing.py:37
# HATPRO channel bandwidths [GHz] (builder-chosen for config 3; the reference is monochromatic)
HATPRO_BW = np.array([0.23] * 11 + [0.6, 1.0, 2.0])


def synth_profiles(P, L=100, seed=1234, top_km=30.0, exponent=2.2, lowest_layer_km=None,
                   clouds=False):
    """Returns dict of (P,L) fp64 arrays, ground-first: z [km], p [mb], t [K], rh [fraction]
    (+ denliq, denice [g m^-3] when clouds=True)."""
    rng = np.random.default_rng(seed)
    s = np.linspace(0.0, 1.0, L)
    if lowest_layer_km is None:
        shape = s ** exponent
    else:  # ICON-D2-like stretching: fixed lowest layer thickness, power law above
        a = lowest_layer_km * (L - 1) / top_km
        shape = a * s + (1.0 - a) * s ** exponent
    z0 = rng.uniform(0.0, 0.5, (P, 1))
    z = z0 + top_km * shape[None, :]
    h = z - z0                                                   # height above ground

    t0 = np.clip(rng.normal(288.0, 8.0, (P, 1)), 255.0, 310.0)
    lapse = 6.5 * rng.normal(1.0, 0.1, (P, 1))
    ztrop = rng.uniform(9.0, 14.0, (P, 1))
    t = t0 - lapse * np.minimum(h, ztrop)
    t = t + np.maximum(h - 20.0, 0.0) * 1.0
    inv = (rng.uniform(0, 1, (P, 1)) < 0.3) * rng.uniform(0.0, 5.0, (P, 1))
    t = t - inv * np.clip(1.0 - h / 0.5, 0.0, 1.0)               # surface-based inversion
    t = np.clip(t, 180.0, 320.0)

    # hydrostatic pressure, layer by layer with the layer-mean temperature
    p0 = rng.normal(1000.0, 15.0, (P, 1))
    tm = 0.5 * (t[:, 1:] + t[:, :-1])
    dlnp = -9.80665 * (h[:, 1:] - h[:, :-1]) * 1000.0 / (287.04 * tm)
    p = p0 * np.exp(np.concatenate([np.zeros((P, 1)), np.cumsum(dlnp, axis=1)], axis=1))

    rh0 = rng.uniform(0.3, 0.95, (P, 1))
    hs = rng.uniform(1.5, 3.0, (P, 1))
    noise = rng.normal(0.0, 0.1, (P, L))
    k = np.array([0.25, 0.5, 0.25])
    noise = np.apply_along_axis(lambda v: np.convolve(v, k, mode="same"), 1, noise)
    rh = np.clip(rh0 * np.exp(-h / hs) + noise, 0.01, 1.0)
    strat = rng.uniform(0.01, 0.05, (P, 1))
    rh = np.where(h > ztrop, strat, rh)

    out = dict(z=np.ascontiguousarray(z), p=np.ascontiguousarray(p), t=np.ascontiguousarray(t),
               rh=np.ascontiguousarray(rh))
    if clouds:
        has = rng.uniform(0, 1, (P, 1)) < 0.5
        base = rng.uniform(0.5, 3.0, (P, 1))
        thick = rng.uniform(0.2, 1.0, (P, 1))
        inside = (h >= base) & (h <= base + thick) & has & (t > 253.0)
        lwc = np.where(inside, np.minimum(0.5, 1.0 * (h - base) + 0.02), 0.0)  # adiabatic-shaped
        out["denliq"] = np.ascontiguousarray(lwc)
        out["denice"] = np.zeros_like(lwc)
        out["rh"] = np.ascontiguousarray(np.where(inside, 1.0, rh))
    return out


def band_frequencies(nsub=50, centres=HATPRO_FRQ, bandwidths=HATPRO_BW):
    """Config 3: nsub equally spaced monochromatic frequencies across each channel's top-hat band.
    Returns (C*nsub,) frequencies, channel-major."""
    off = (np.arange(nsub) + 0.5) / nsub - 0.5
    return (np.asarray(centres)[:, None] + np.asarray(bandwidths)[:, None] * off[None, :]).reshape(-1)


def dense_spectrum(f0=20.0, f1=60.0, step=0.01):
    """Config 4: 20-60 GHz at 10 MHz steps (4001 frequencies)."""
    n = int(round((f1 - f0) / step)) + 1
    return f0 + step * np.arange(n)
