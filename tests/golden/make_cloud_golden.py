#!/usr/bin/env python3
"""Regenerate tests/golden/cloud_reference.npz -- outputs of the REFERENCE's OWN
`derive_cloud_features` (/root/reference/python_src/preproc/derive_cloud_water.py:146-363, with
`calc_lwc` :68-142) on seeded synthetic 180-level profiles (top-first, like `Level_*`).
The reference file is imported unmodified; xarray / matplotlib / pyrtlib, which it imports but does not
use in these functions, are stubbed.  `/root/reference` is read only HERE; tests read the .npz.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))


def _stubs():
    for name in ("xarray", "matplotlib", "matplotlib.pyplot", "pyrtlib", "pyrtlib.climatology", "pyrtlib.utils"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].use = lambda *a, **k: None
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["pyrtlib.climatology"].AtmosphericProfiles = object
    sys.modules["pyrtlib.utils"].ppmv2gkg = None
    sys.modules["pyrtlib.utils"].mr2rh = None


def synth(rng, n=180, layers=2, warm=False):
    """Top-first column: z from ~30 km down to ~100 m (dense near the ground), T, p, RH with
    `layers` moist layers placed in different altitude bands."""
    s = np.linspace(1.0, 0.0, n)
    z = 100.0 + 30000.0 * s ** 2.2 + rng.uniform(0, 30)
    t0 = rng.uniform(285, 303) if warm else rng.uniform(262, 296)
    t = np.where(z < 11000, t0 - 6.5e-3 * z, t0 - 71.5 + 1e-3 * (z - 11000)) + rng.normal(0, 0.3, n)
    p = 1013.0 * np.exp(-z / 7600.0)
    rh = np.clip(rng.uniform(30, 70) * np.exp(-z / rng.uniform(3000, 9000)) + rng.normal(0, 3, n), 1, 100)
    for _ in range(layers):
        zc = rng.choice([rng.uniform(300, 1800), rng.uniform(2200, 5500), rng.uniform(6500, 11000),
                         rng.uniform(12500, 15000)])
        th = rng.uniform(60, 1500)
        peak = rng.uniform(86, 101)
        m = np.abs(z - zc) < th / 2
        rh[m] = np.maximum(rh[m], peak + rng.normal(0, 1.5, m.sum()))
    return p, t, z, np.clip(rh, 1, 105)


def main():
    _stubs()
    sys.path.insert(0, "/root/reference/python_src/preproc")
    import derive_cloud_water as ref   # noqa: E402  (the reference file itself)
    rng = np.random.default_rng(20261019)
    P, T, Z, R, out = [], [], [], [], []
    for c in range(160):
        p, t, z, rh = synth(rng, layers=int(rng.integers(0, 5)), warm=bool(c % 3 == 0))
        with np.errstate(all="ignore"):
            r = ref.derive_cloud_features(p, t, None, None, z, rh)
        P.append(p); T.append(t); Z.append(z); R.append(rh)
        out.append(r)
    np.savez_compressed(os.path.join(HERE, "cloud_reference.npz"),
                        p=np.array(P), t=np.array(T), z=np.array(Z), rh=np.array(R),
                        lwc_m3=np.array([o[0] for o in out]), lwc_kg=np.array([o[1] for o in out]),
                        lwp=np.array([o[2] for o in out]), iwc_m3=np.array([o[3] for o in out]),
                        iwc_kg=np.array([o[4] for o in out]), iwp=np.array([o[5] for o in out]))
    print("cases with liquid:", int(sum((o[0] > 0).any() for o in out)), "with ice:",
          int(sum((o[3] > 0).any() for o in out)), "of", len(out))


if __name__ == "__main__":
    main()
