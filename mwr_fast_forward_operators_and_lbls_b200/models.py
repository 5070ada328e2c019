"""Absorption-model handles: the host side of `TbCloudRTE.init_absmdl(mdl)`
(`/root/reference/python_src/proc/PyRTlib_processing.py:124,132,140,148`).

A model is an immutable device-resident table set (`mwrt_model*`).  Built-in models come from the
tables compiled into libmwrt.so (generated from `data/models/*.json` by tools/gen_model_tables.py --
UNVERIFIED recollection of the PyRTlib 1.1.x line lists); `Model.from_descriptor` uploads any
descriptor of the same schema, e.g. the true tables dumped by `baseline/probe_pyrtlib.py`.
"""
import ctypes as C
import json
import os
import warnings

import numpy as np

from . import _lib

DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "models")

# what the reference declares (PyRTlib_processing.py:90) and runs (:122,130,138,146)
REFERENCE_MODELS = ("R17", "R03", "R16", "R19", "R98", "R19SD", "R20", "R20SD", "R24")


def implemented_models():
    """Model strings accepted by `init_absmdl` (PyRTlib: AbsModel.implemented_models())."""
    lib = _lib.load()
    return [lib.mwrt_builtin_model_name(i).decode() for i in range(lib.mwrt_builtin_model_count())]


def load_descriptor(name):
    """The JSON descriptor of a built-in model (for inspection / perturbation in tests)."""
    path = os.path.join(DATA_DIR, name + ".json")
    if not os.path.exists(path):
        raise ValueError("absorption model '%s' is not implemented" % name)
    with open(path) as fh:
        return json.load(fh)


def _pack(desc):
    h, o, n2 = desc["h2o"], desc["o2"], desc["n2"]
    h2o = np.ascontiguousarray([h[c] for c in _lib.H2O_COLS], dtype=np.float64)
    o2 = np.ascontiguousarray([o[c] for c in _lib.O2_COLS], dtype=np.float64)
    d = _lib.ModelDesc()
    d.n_h2o, d.n_o2 = h2o.shape[1], o2.shape[1]
    d.h2o_sd = int(bool(h["sd"]))
    d.o2_mixing = int(o["mixing"])
    d.o2_dens118 = int(bool(o["dens118"]))
    d.o2_sd118 = int(bool(o["sd118"]))
    d.n2_fdep = int(bool(n2["fdep"]))
    d.liq_model = 0 if desc["liq"]["model"] == "liebe91" else 1
    for k in ("reftline", "reftcon", "cf", "xcf", "cs", "xcs", "pvap_div", "den_const", "line_const"):
        setattr(d, "h2o_" + k, float(h[k]))
    for k in ("x", "wb300", "pvap_div", "h2o_broad", "nonres", "abs_num", "abs_den", "iso"):
        setattr(d, "o2_" + k, float(o[k]))
    d.n2_l, d.n2_m, d.n2_n = float(n2["l"]), float(n2["m"]), float(n2["n"])
    d.h2o_lines = h2o.ctypes.data_as(_lib.c_dp)
    d.o2_lines = o2.ctypes.data_as(_lib.c_dp)
    return d, (h2o, o2)  # keep the arrays alive while the C call reads them


# Provenance of the built-in tables.  They answer to PyRTlib's model strings but were restated without
# access to PyRTlib (no network, package absent: DESIGN.md section 2): nothing pins their digits to the
# upstream line lists, so every user of a built-in model is told once, and the driver records it in the
# output attributes.  `Model.from_descriptor` with tables dumped from a real PyRTlib carries no flag.
BUILTIN_TABLES_VERIFIED = False
PROVENANCE = ("built-in %s tables: restatement of the PyRTlib 1.1.x / Rosenkranz line lists, NOT verified "
              "against PyRTlib output (mwr_fast_forward_operators_and_lbls_b200/DESIGN.md, 'Oracle')")
_warned = set()


def provenance(name):
    """One-line provenance note for a model name ('' for descriptors supplied by the caller)."""
    return PROVENANCE % name if not BUILTIN_TABLES_VERIFIED else ""


class UnverifiedModelTablesWarning(UserWarning):
    pass


class Model:
    """Owns one `mwrt_model*`.  Immutable after creation, safe to share between plans."""

    def __init__(self, name):
        lib = _lib.load()
        self.name = name
        self.verified = BUILTIN_TABLES_VERIFIED
        if not self.verified and name not in _warned and not os.environ.get("MWRT_QUIET"):
            _warned.add(name)
            warnings.warn(provenance(name), UnverifiedModelTablesWarning, stacklevel=3)
        h = C.c_void_p()
        _lib.check(lib.mwrt_model_create(name.encode(), C.byref(h)), "init_absmdl('%s')" % name)
        self._h = h

    @classmethod
    def from_descriptor(cls, desc):
        lib = _lib.load()
        d, keep = _pack(desc)
        self = cls.__new__(cls)
        self.name = desc.get("name", "custom")
        self.verified = bool(desc.get("verified", False))
        h = C.c_void_p()
        _lib.check(lib.mwrt_model_create_custom(C.byref(d), C.byref(h)), "model_create_custom")
        del keep
        self._h = h
        return self

    @property
    def handle(self):
        return self._h

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.load().mwrt_model_destroy(h)
            except Exception:
                pass


_cache = {}


def get_model(model):
    """Model by name (cached per process) or pass a Model through."""
    if isinstance(model, Model):
        return model
    if isinstance(model, dict):
        return Model.from_descriptor(model)
    import torch
    dev = torch.cuda.current_device() if torch.cuda.is_available() else 0
    _lib.set_device(dev)
    if (model, dev) not in _cache:
        _cache[(model, dev)] = Model(model)
    return _cache[(model, dev)]
