"""Drop-in for `pyrtlib.tb_spectrum.TbCloudRTE` on the reference's call pattern:

    rte = TbCloudRTE(z_in[::-1], p_in[::-1], t_in[::-1], rh_in[::-1], frqs, ang)
    rte.init_absmdl(mdl)
    rte.satellite = False
    df = rte.execute()
    tbs[i, :, k, j] = df["tbtotal"].values

(`/root/reference/python_src/proc/PyRTlib_processing.py:123-127`, repeated at :131-151 and in
`old_processing.py:556-565`).  Swapping the import line is the whole migration.  The class is a
P=1 shim over the batched plan (`forward.Plan`); all arithmetic runs in libmwrt.so on the GPU.

Same names, argument meaning and error behaviour as upstream:
unknown model -> ValueError; negative absorption in the layer integration -> ValueError;
inputs are never modified.  Unlike upstream, all angles share one absorption computation.
"""
import threading

import numpy as np
import pandas as pd
import torch

from . import _lib
from .forward import FIELDS, Plan, _dev64
from .models import get_model, implemented_models

COLUMNS = ("tbtotal", "tbatm", "tmr", "tmrcld", "tauwet", "taudry", "tauliq", "tauice")


class TbCloudRTE:
    """z [km], p [mb], t [K], rh [fraction], frq [GHz], angles [deg elevation]; ground-first."""

    def __init__(self, z, p, t, rh, frq, angles, o3n=None, amu=None, absmdl="", ray_tracing=False,
                 from_sat=True, cloudy=False):
        self.z = np.array(z, dtype=np.float64)   # copies: the reference reuses its arrays (:123-147)
        self.p = np.array(p, dtype=np.float64)
        self.tk = np.array(t, dtype=np.float64)
        self.rh = np.array(rh, dtype=np.float64)
        self.frq = np.atleast_1d(np.array(frq, dtype=np.float64))
        self.angles = np.atleast_1d(np.array(angles, dtype=np.float64))
        if not (self.z.shape == self.p.shape == self.tk.shape == self.rh.shape and self.z.ndim == 1):
            raise ValueError("z, p, t, rh must be 1-D arrays of the same length")
        if o3n is not None:
            raise NotImplementedError("ozone absorption (o3n) is outside the reference's path")
        if amu is not None:
            raise NotImplementedError("spectroscopic-uncertainty perturbation (amu) is outside the "
                                      "reference's path")
        self.ray_tracing = bool(ray_tracing)
        self.satellite = bool(from_sat)
        self.cloudy = bool(cloudy)
        self.emissivity = 1.0
        self.nl, self.nf, self.nang = len(self.z), len(self.frq), len(self.angles)
        self.denliq = np.zeros(self.nl)
        self.denice = np.zeros(self.nl)
        self.cldh = np.zeros((2, 0))
        self._model = None
        if absmdl:
            self.init_absmdl(absmdl)

    @staticmethod
    def implemented_models():
        return implemented_models()

    def init_absmdl(self, absmdl):
        """Select the absorption model; ValueError for an unknown string (upstream behaviour)."""
        self._model = get_model(absmdl)
        self.absmdl = absmdl

    def init_cloudy(self, cldh, denice, denliq):
        """cldh: (2, ncld) cloud base/top heights [km]; denice/denliq [g m^-3] per level."""
        self.cloudy = True
        self.cldh = np.array(cldh, dtype=np.float64).reshape(2, -1)
        self.denice = np.array(denice, dtype=np.float64)
        self.denliq = np.array(denliq, dtype=np.float64)
        if self.denice.shape != self.z.shape or self.denliq.shape != self.z.shape:
            raise ValueError("denice/denliq must have one value per level")

    def _cloud_levels(self):
        """Level indices of the first cloud layer's base and top (`cldh` heights are matched to the
        nearest level, heights taken above the antenna like upstream's `z -= z[0]`)."""
        if not (self.cloudy and self.cldh.shape[1] > 0) or self.satellite:
            return None
        zrel = self.z - self.z[0]
        base, top = self.cldh[:, 0] - self.z[0]
        return int(np.argmin(np.abs(zrel - base))), int(np.argmin(np.abs(zrel - top)))

    def execute(self, only_bt=True):
        if self._model is None:
            raise ValueError("init_absmdl() must be called before execute()")
        dl = di = None
        if self.cloudy:
            dl, di = _dev64(self.denliq[None]), _dev64(self.denice[None])
        cld = self._cloud_levels()
        names = FIELDS + (("tmrcld",) if cld is not None else ())
        # one device buffer for all fields -> one device-to-host copy
        buf = torch.empty((len(names), 1, self.nf, self.nang), dtype=torch.float64, device="cuda")
        out = {f: buf[i] for i, f in enumerate(names)}
        kw = {}
        if cld is not None:
            kw = dict(cld_ibase=torch.tensor([cld[0]], dtype=torch.int32, device="cuda"),
                      cld_itop=torch.tensor([cld[1]], dtype=torch.int32, device="cuda"))
        with _PLAN_LOCK:                  # cached plans are shared; a plan runs one call at a time
            plan = _cached_plan(self._model, self.nl, self.frq, self.angles, self.ray_tracing,
                                self.satellite, self.cloudy, self.emissivity)
            res = plan.forward(_dev64(self.z[None]), _dev64(self.p[None]), _dev64(self.tk[None]),
                               _dev64(self.rh[None]), dl, di, fields=names, out=out, **kw)
            host = buf.cpu().numpy()      # synchronises the stream
            status = int(res["status"][0].item())
        if status & _lib.ST_NEG_ABS:
            raise ValueError("Error encountered in exponential_integration")  # upstream's message
        cols = {}
        for c in COLUMNS:
            if c in names:
                cols[c] = host[names.index(c), 0].T.reshape(self.nf * self.nang)  # angle-major
            else:      # tmrcld without a cloud layer: 0, as upstream
                cols[c] = np.zeros(self.nf * self.nang)
        df = pd.DataFrame(cols)
        df["angle"] = np.repeat(self.angles, self.nf)
        df["frq"] = np.tile(self.frq, self.nang)
        self.status = status
        return df


# The reference builds a fresh TbCloudRTE for every (profile, elevation, model)
# (PyRTlib_processing.py:99-151: 41 680 objects per run, but only 10 elevations x 4 models x 1 level
# count of distinct configurations).  Plans -- device copies of frq/angles, workspace, streams -- are
# therefore cached per configuration; a plan is used by one execute() at a time.
_PLAN_CACHE = {}
_PLAN_CACHE_MAX = 64
_PLAN_LOCK = threading.Lock()


def _cached_plan(model, nl, frq, angles, ray_tracing, from_sat, cloudy, emissivity):
    key = (model.name if hasattr(model, "name") else id(model), id(model), int(nl), frq.tobytes(),
           angles.tobytes(), bool(ray_tracing), bool(from_sat), bool(cloudy), float(emissivity),
           torch.cuda.current_device())
    plan = _PLAN_CACHE.get(key)
    if plan is None:
        if len(_PLAN_CACHE) >= _PLAN_CACHE_MAX:
            _PLAN_CACHE.pop(next(iter(_PLAN_CACHE)))          # oldest entry
        plan = Plan(model, nl, frq, angles, ray_tracing=ray_tracing, from_sat=from_sat, cloudy=cloudy,
                    tau_layers=False, emissivity=emissivity)
        _PLAN_CACHE[key] = plan
    return plan
