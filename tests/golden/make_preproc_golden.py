#!/usr/bin/env python3
"""Regenerate tests/golden/preproc_reference.npz -- outputs of the REFERENCE's OWN pre-processing
functions (N1) on seeded synthetic sondes.

Unlike the RTE goldens these ARE reference output: `/root/reference/python_src/preproc/
preprocessing4all.py` is imported in this container and its `read_radiosonde_nc_arms` (:192-343,
thinning :282-285, running means :156-169, humidity :104-130, jump check :315-331) and
`add_clim2profiles` (:478-531, incl. `lowest2tenhPa` :450-474) are run unmodified, followed by the
lowest-180 selection of :807-814.  Modules the file imports but this container lacks (xarray,
matplotlib, pyrtlib, the repo's own MWR readers) are replaced by stubs:
  * `xr.open_dataset` returns an in-memory fake holding the synthetic sonde (no file I/O semantics);
  * `pyrtlib.climatology.AtmosphericProfiles.gl_atm` / `ppmv2gkg` / `mr2rh` return the climatology
    passed in (pyrtlib itself is absent -- the climatology digits are the only unpinned input, and
    they are an explicit argument of the product and of the oracle).
`/root/reference` is only read HERE, at generation time; the tests read the committed .npz.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
REF = "/root/reference/python_src/preproc"


class _Var:
    def __init__(self, values):
        self.values = np.asarray(values)

    def isel(self, **kw):
        (idx,) = kw.values()
        return _Var(self.values[idx])


class _FakeDS(dict):
    @property
    def data_vars(self):
        return self.keys()


_CURRENT = {}
_CLIM = {}


def _install_stubs():
    xr = types.ModuleType("xarray")
    xr.open_dataset = lambda f, *a, **k: _CURRENT[f]
    sys.modules["xarray"] = xr
    for name in ("matplotlib", "matplotlib.pyplot", "derive_cloud_water", "MWR_read_in_module",
                 "pyrtlib", "pyrtlib.climatology", "pyrtlib.utils"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].use = lambda *a, **k: None
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["derive_cloud_water"].derive_cloud_features = None
    sys.modules["MWR_read_in_module"].get_mwr_data = None

    class AtmosphericProfiles:
        H2O = 0

        @staticmethod
        def gl_atm(atm=1):
            assert atm == 1
            c = _CLIM
            return c["z_km"].copy(), c["p"].copy(), None, c["t"].copy(), c["gkg"][:, None].copy()
    sys.modules["pyrtlib.climatology"].AtmosphericProfiles = AtmosphericProfiles
    # the stubbed gl_atm hands over g/kg directly, so ppmv2gkg is the identity and mr2rh returns the
    # climatology's RH column
    sys.modules["pyrtlib.utils"].ppmv2gkg = lambda x, gas: x
    sys.modules["pyrtlib.utils"].mr2rh = lambda p, t, gkg: (_CLIM["rh"].copy(), None)


def synth_sonde(rng, n=3200, frac_rh=False, jump=False, top_km=30.0, nan_frac=0.01, hold=3):
    """Raw 1 Hz-like sonde: monotone ascent to `top_km`, hydrostatic p, lapse-rate T, decaying RH."""
    h0 = rng.uniform(50, 150)
    dz = np.full(n, top_km * 1000 / n) * rng.uniform(0.8, 1.2, n)
    h = h0 + np.concatenate([[0.0], np.cumsum(dz[:-1])])
    h[:hold] = h0 + np.arange(hold) * 0.5         # launch hold: |dh| < 2 m (crop detection :229-241)
    t0 = rng.uniform(278, 300)
    t = np.where(h < 11000, t0 - 6.5e-3 * (h - h0), t0 - 6.5e-3 * (11000 - h0) + 1e-3 * (h - 11000))
    t = t + rng.normal(0, 0.2, n)
    p = rng.uniform(990, 1020) * np.exp(-(h - h0) / 7400.0)
    rh = np.clip(rng.uniform(40, 95) * np.exp(-(h - h0) / rng.uniform(2000, 4000))
                 + rng.normal(0, 2, n), 1.0, 100.0)
    for a in (t, p, rh):
        a[rng.uniform(size=n) < nan_frac] = np.nan
    a = p
    a[0] = a[0] if np.isfinite(a[0]) else 1000.0
    if frac_rh:
        rh = rh / 100.0
    if jump:
        h[200:] += 5000.0
    return h, t, p, rh


def main():
    _install_stubs()
    sys.path.insert(0, REF)
    import preprocessing4all as ref   # noqa: E402  (the reference file itself, unmodified)
    from mwr_fast_forward_operators_and_lbls_b200 import preproc as prod   # climatology tables only
    _CLIM.update(prod.midlatitude_summer())
    rng = np.random.default_rng(20261018)
    cases = [dict(), dict(frac_rh=True), dict(jump=True), dict(top_km=12.5), dict(n=5200, top_km=34.0),
             dict(n=900, top_km=26.0), dict(nan_frac=0.08), dict(hold=1), dict(n=2400, top_km=22.0),
             dict(n=4000, top_km=16.0)]
    out = {}
    meta = []
    for c, kw in enumerate(cases):
        h, t, p, rh = synth_sonde(rng, **kw)
        for vital in (False, True):
            fname = ("/x/Vital_I/radiosondes/%d.nc" if vital else "/x/other/%d.nc") % c
            _CURRENT[fname] = _FakeDS(Height=_Var(h), Temperature=_Var(t), Pressure=_Var(p),
                                      Humidity=_Var(rh), Latitude=_Var([50.9]), Longitude=_Var([6.4]))
            r = ref.read_radiosonde_nc_arms(file=fname, crop=0)
            length_value, p_a, t_a, ppmv_a, _, _, m_a, z_a, rh_a, _ = r
            key = "c%d%s" % (c, "v" if vital else "")
            rejected = length_value == 0
            if not rejected:
                p2, t2, q2, m2, z2, r2 = ref.add_clim2profiles(np.array(p_a), np.array(t_a),
                                                               np.array(ppmv_a), np.array(m_a),
                                                               np.array(z_a), np.array(rh_a))
                n = ref.n_levels
                if len(p2) < n:
                    rejected = True      # the reference would raise on the (180,) assignment :807
                else:
                    out[key + "_levels"] = np.stack([p2[-n:], t2[-n:], q2[-n:], 1000 * m2[-n:],
                                                     z2[-n:], r2[-n:]])
                out[key + "_thinned"] = np.stack([z_a, t_a, p_a, rh_a])
            if not vital:
                out[key + "_raw"] = np.stack([h, t, p, rh])
            meta.append((key, int(rejected)))
    out["keys"] = np.array([k for k, _ in meta])
    out["rejected"] = np.array([r for _, r in meta])
    for k in ("z_km", "p", "t", "gkg", "rh"):
        out["clim_" + k] = _CLIM[k]
    np.savez_compressed(os.path.join(HERE, "preproc_reference.npz"), **out)
    print("cases:", meta)


if __name__ == "__main__":
    main()
