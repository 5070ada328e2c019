"""N2: the wire / on-disk formats either side of the path.

* NetCDF variable contract in: `Level_z` [m], `Level_Pressure` [hPa], `Level_Temperature` [K], `Level_RH` [%]
  with dims (N_Levels, time, Crop), index 0 = top (`preprocessing4all.py:1195-1203`, `:530-531`),
  plus the `elevation` coordinate (`PyRTlib_processing.py:101`).
* NetCDF variable contract out: `TBs_PyRTlib_{R24,R17,R98,R20}` with dims
  ('time','N_Channels','elevation','Crop') and the attrs of `PyRTlib_processing.py:161-195`; the merge
  step copies exactly these names (`summarize_proc_results.py:84-87`).
* Legacy CSV (`<sonde>_out.txt`): a `tbtotal` column of 252 rows = {cropped, uncropped} x 9 models
  [R17,R03,R16,R19,R98,R19SD,R20,R20SD,R24] x 14 channels, the slicing of `old_merge2nc.py:361-362,417-435`.

xarray / netCDF4 are not installed in the build image: `open_dataset` / `SimpleDataset.to_netcdf` use
xarray when it is importable (NETCDF4_CLASSIC like the reference, `PyRTlib_processing.py:211`) and fall
back to scipy's NetCDF-3 reader/writer otherwise.
"""
import numpy as np
import pandas as pd

LEGACY_MODELS = ("R17", "R03", "R16", "R19", "R98", "R19SD", "R20", "R20SD", "R24")
N_CHANNELS = 14
OUTPUT_VARS = (("R24", "TBs_PyRTlib_R24"), ("R17", "TBs_PyRTlib_R17"), ("R98", "TBs_PyRTlib_R98"),
               ("R20", "TBs_PyRTlib_R20"))
TB_DIMS = ("time", "N_Channels", "elevation", "Crop")


def tb_attrs(model):
    """Attributes the reference attaches to each TBs_PyRTlib_* variable (PyRTlib_processing.py:162-168)."""
    return {"long_name": "Brightness temperature modelled by %s" % model, "units": "K",
            "standard_name": "brightness_temperature",
            "comments": "Brightness temperatures modeled from radiosonde data for 14 channels of "
                        "HATPRO radiometer"}


# ---------------------------------------------------------------------------------------------
# legacy CSV
# ---------------------------------------------------------------------------------------------
def write_legacy_csv(path, tb_cropped, tb_uncropped, extra_columns=None):
    """tb_cropped / tb_uncropped: dict model -> (14,) zenith tbtotal.  Writes the 252-row CSV."""
    rows = []
    for block in (tb_cropped, tb_uncropped):
        for m in LEGACY_MODELS:
            v = np.asarray(block[m], dtype=np.float64).reshape(-1)
            if v.shape != (N_CHANNELS,):
                raise ValueError("model %s: expected %d channel values" % (m, N_CHANNELS))
            rows.append(v)
    df = pd.DataFrame({"tbtotal": np.concatenate(rows)})
    for k, v in (extra_columns or {}).items():
        df[k] = v
    df.to_csv(path, index=False)
    return df


def read_legacy_csv(path):
    """Inverse of the slicing at old_merge2nc.py:417-435: returns (cropped, uncropped) dicts."""
    tb = pd.read_csv(path)["tbtotal"].values
    if len(tb) != 2 * len(LEGACY_MODELS) * N_CHANNELS:
        raise ValueError("expected %d rows, found %d" % (2 * len(LEGACY_MODELS) * N_CHANNELS, len(tb)))
    out = []
    for b in range(2):
        base = b * len(LEGACY_MODELS) * N_CHANNELS
        out.append({m: tb[base + i * N_CHANNELS: base + (i + 1) * N_CHANNELS]
                    for i, m in enumerate(LEGACY_MODELS)})
    return out[0], out[1]


# ---------------------------------------------------------------------------------------------
# minimal dataset (xarray-shaped) for environments without xarray
# ---------------------------------------------------------------------------------------------
class Variable:
    def __init__(self, dims, values, attrs=None):
        self.dims = tuple(dims)
        self.values = np.asarray(values)
        self.attrs = dict(attrs or {})
        if self.values.ndim != len(self.dims):
            raise ValueError("dims %r do not match array of rank %d" % (self.dims, self.values.ndim))


class SimpleDataset:
    """The few xarray.Dataset behaviours the reference's driver touches: `ds[name].values`,
    `ds[name] = (dims, array)`, `.attrs`, `.to_netcdf`."""

    def __init__(self, variables=None):
        self.variables = {}
        for k, v in (variables or {}).items():
            self[k] = v

    def __getitem__(self, name):
        return self.variables[name]

    def __setitem__(self, name, value):
        if isinstance(value, Variable):
            self.variables[name] = value
        else:
            dims, arr = value[0], value[1]
            self.variables[name] = Variable(dims, arr, value[2] if len(value) > 2 else None)

    def __contains__(self, name):
        return name in self.variables

    def keys(self):
        return self.variables.keys()

    def to_netcdf(self, path, format="NETCDF3_64BIT"):
        from scipy.io import netcdf_file
        with netcdf_file(path, "w", version=2) as nc:
            dims = {}
            for v in self.variables.values():
                for d, n in zip(v.dims, v.values.shape):
                    if d in dims and dims[d] != n:
                        raise ValueError("inconsistent size for dimension %s" % d)
                    dims[d] = n
            for d, n in dims.items():
                nc.createDimension(d, n)
            for name, v in self.variables.items():
                arr = v.values
                if arr.dtype == np.bool_:
                    arr = arr.astype(np.int8)
                if arr.dtype == np.int64:
                    arr = arr.astype(np.int32)
                var = nc.createVariable(name, arr.dtype, v.dims)
                var[:] = arr
                for ak, av in v.attrs.items():
                    setattr(var, ak, av)


def open_dataset(path):
    """xarray.open_dataset when xarray is installed (any NetCDF flavour); otherwise a SimpleDataset
    read with scipy (NetCDF-3 classic / 64-bit offset only)."""
    try:
        import xarray as xr  # noqa: F401
        return xr.open_dataset(path)
    except ImportError:
        pass
    from scipy.io import netcdf_file
    with open(path, "rb") as fh:
        if fh.read(4)[:3] != b"CDF":
            raise RuntimeError("%s is not NetCDF-3; reading NetCDF-4/HDF5 needs xarray + netCDF4, which "
                               "are not installed" % path)
    ds = SimpleDataset()
    with netcdf_file(path, "r", mmap=False) as nc:
        for name, var in nc.variables.items():
            attrs = {k: (v.decode() if isinstance(v, bytes) else v) for k, v in var._attributes.items()}
            ds[name] = Variable(var.dimensions, np.array(var[:]), attrs)
    return ds
