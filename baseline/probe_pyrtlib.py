#!/usr/bin/env python3
"""If real `pyrtlib` is importable on this box, record TRUE goldens for BASELINE.json config 1 in the
schema of tests/golden/config1_us_standard.json (and say which version).  It is not importable in the
build container, so this is a probe: exit 0 with {"pyrtlib": null} otherwise.

    python baseline/probe_pyrtlib.py [out.json]
"""
import json
import sys

FRQ = [22.24, 23.04, 23.84, 25.44, 26.24, 27.84, 31.4, 51.26, 52.28, 53.86, 54.94, 56.66, 57.3, 58.0]
ANG = [90., 30, 19.2, 14.4, 11.4, 8.4, 6.6, 5.4, 4.8, 4.2]
MODELS = ["R98", "R17", "R20", "R24", "R22SD", "R19SD"]


def main():
    try:
        import pyrtlib
        from pyrtlib.climatology import AtmosphericProfiles as atmp
        from pyrtlib.tb_spectrum import TbCloudRTE
        from pyrtlib.utils import ppmv2gkg, mr2rh
    except Exception as exc:  # noqa: BLE001
        print(json.dumps({"pyrtlib": None, "why": repr(exc)}))
        return 0
    import numpy as np
    z, p, _, t, md = atmp.gl_atm(atmp.US_STANDARD)
    gkg = ppmv2gkg(md[:, atmp.H2O], atmp.H2O)
    rh = mr2rh(p, t, gkg)[0] / 100
    out = dict(source="pyrtlib %s" % getattr(pyrtlib, "__version__", "?"), z=z.tolist(), p=p.tolist(),
               t=t.tolist(), rh=rh.tolist(), frq=FRQ, angles=ANG, cases=[])
    for mdl in MODELS:
        for rt in (False, True):
            try:
                rte = TbCloudRTE(z, p, t, rh, np.array(FRQ), np.array(ANG), ray_tracing=rt)
                rte.init_absmdl(mdl)
                rte.satellite = False
                df = rte.execute()
            except Exception as exc:  # noqa: BLE001
                out["cases"].append(dict(model=mdl, ray_tracing=rt, error=repr(exc)))
                continue
            out["cases"].append(dict(model=mdl, ray_tracing=rt,
                                     **{k: df[k].values.tolist() for k in ("tbtotal", "tbatm", "tmr",
                                                                           "tauwet", "taudry")}))
    path = sys.argv[1] if len(sys.argv) > 1 else "pyrtlib_goldens.json"
    with open(path, "w") as fh:
        json.dump(out, fh)
    print(json.dumps({"pyrtlib": out["source"], "wrote": path, "cases": len(out["cases"])}))
    return 0


if __name__ == "__main__":
    sys.exit(main())
