#!/usr/bin/env python3
"""Regenerate tests/golden/config1_us_standard.json.

NOT true PyRTlib goldens: the reference's engine (pyrtlib==1.1.1) is absent and not installable
offline, so these vectors come from the in-repo scalar restatement
(oracle/pyrtlib_scalar.py) and pin it against regressions -- "parity unpinned".  If a box has pyrtlib,
`baseline/probe_pyrtlib.py` writes the same schema from the real package.

BASELINE.json config 1: AFGL US-standard atmosphere, 14 HATPRO channels, 10 elevations, clear sky,
downwelling (`satellite=False`), the reference's models (PyRTlib_processing.py:122-151) + R22SD.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import atmospheres, pyrtlib_scalar  # noqa: E402

FRQ = [22.24, 23.04, 23.84, 25.44, 26.24, 27.84, 31.4, 51.26, 52.28, 53.86, 54.94, 56.66, 57.3, 58.0]
ANG = [90., 30, 19.2, 14.4, 11.4, 8.4, 6.6, 5.4, 4.8, 4.2]
MODELS = ["R98", "R17", "R20", "R24", "R22SD", "R19SD"]


def main():
    z, p, t, rh = atmospheres.us_standard()
    out = dict(source="oracle/pyrtlib_scalar.py (in-repo restatement; NOT PyRTlib output)",
               z=z.tolist(), p=p.tolist(), t=t.tolist(), rh=rh.tolist(), frq=FRQ, angles=ANG, cases=[])
    for mdl in MODELS:
        for rt in (False, True):
            rte = pyrtlib_scalar.TbCloudRTE(z, p, t, rh, np.array(FRQ), np.array(ANG), ray_tracing=rt)
            rte.init_absmdl(mdl)
            rte.satellite = False
            d = rte.execute()
            out["cases"].append(dict(model=mdl, ray_tracing=rt,
                                     **{k: d[k].tolist() for k in ("tbtotal", "tbatm", "tmr", "tauwet",
                                                                   "taudry")}))
    with open(os.path.join(HERE, "config1_us_standard.json"), "w") as fh:
        json.dump(out, fh)
    print("wrote", len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
