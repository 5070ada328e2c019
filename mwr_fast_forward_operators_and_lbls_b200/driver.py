#!/usr/bin/env python3
"""Batched drop-in for the reference's driver `derive_TBs4PyRTlib`
(`/root/reference/python_src/proc/PyRTlib_processing.py:83-197`) and its CLI (`:43-65,203-211`), plus the
legacy per-sonde runner's output (`old_run_pyrtlib/run_pyrtlib_on_all.py:45-58` -> `*_out.txt`).

The reference loops time x Crop x elevation in Python and builds one `TbCloudRTE` per iteration and
model (41 680 `execute()` calls for 521 sondes); here every (time, Crop) profile, all elevations and
all channels of one model are ONE batched GPU call, i.e. four calls in total.

    python -m mwr_fast_forward_operators_and_lbls_b200.driver -i in.nc -o out.nc
"""
import argparse
import os

import numpy as np

from . import io_formats
from .forward import simulate
from .models import provenance
from .synth import HATPRO_FRQ

ELEVATIONS = np.array([90., 30, 19.2, 14.4, 11.4, 8.4, 6.6, 5.4, 4.8, 4.2])   # PyRTlib_processing.py:37
N_LEVELS = 180                                                               # PyRTlib_processing.py:35


R_L = 287.06   # J kg^-1 K^-1, the air gas constant derive_cloud_water.py:76 converts LWC with


def water_content_to_density(level_kg_per_kg, level_p_hpa, level_t_k):
    """N3 wiring: `Level_Liquid` / `Level_Ice` are stored as kg/kg = LWC / rho_air with
    rho_air = p*100 / (R_L*T) (derive_cloud_water.py:96,104; preprocessing4all.py:811-812).
    PyRTlib's `init_cloudy` wants g m^-3: denliq = Level_Liquid * rho_air * 1000."""
    rho = np.asarray(level_p_hpa, dtype=np.float64) * 100.0 / R_L / np.asarray(level_t_k, dtype=np.float64)
    return np.asarray(level_kg_per_kg, dtype=np.float64) * rho * 1000.0


def derive_tbs_arrays(level_z_m, level_p_hpa, level_t_k, level_rh_pct, elevations=ELEVATIONS,
                      frqs=HATPRO_FRQ, models=("R20", "R24", "R17", "R98"), ray_tracing=False,
                      level_liquid=None, level_ice=None):
    """Level_* arrays of shape (N_Levels, time, Crop), index 0 = top (the NetCDF contract).
    Returns dict model -> (time, n_channels, n_elevations, Crop) tbtotal [K]; NaN wherever the
    reference's `check_for_nans` gate would have skipped the call (PyRTlib_processing.py:71-79,117-119).
    level_liquid / level_ice (kg/kg, the `Level_Liquid` / `Level_Ice` variables) switch on the cloudy
    path; the reference's PyRTlib leg never passes them (clear sky only, old_processing.py:559-562)."""
    z = np.asarray(level_z_m, dtype=np.float64)
    L, T, C = z.shape
    frqs = np.asarray(frqs, dtype=np.float64)
    elevations = np.asarray(elevations, dtype=np.float64)

    def stack(v, div):   # (L,T,C) top-first -> (T*C, L) ground-first; units exactly as :109-114
        v = np.asarray(v, dtype=np.float64) / div
        return np.ascontiguousarray(v[::-1].reshape(L, T * C).T)
    zk, p, t, rh = stack(z, 1000.0), stack(level_p_hpa, 1.0), stack(level_t_k, 1.0), stack(level_rh_pct, 100.0)
    out = {m: np.full((T, len(frqs), len(elevations), C), np.nan) for m in models}
    good = ~np.isnan(elevations)
    if np.isnan(frqs).any() or not good.any():
        return out                                    # the reference would skip every call
    denliq = denice = None
    if level_liquid is not None or level_ice is not None:
        zero = np.zeros_like(z)
        denliq = stack(water_content_to_density(zero if level_liquid is None else level_liquid,
                                                level_p_hpa, level_t_k), 1.0)
        denice = stack(water_content_to_density(zero if level_ice is None else level_ice,
                                                level_p_hpa, level_t_k), 1.0)
    for m in models:
        res = simulate(zk, p, t, rh, frqs, elevations[good], m, ray_tracing=ray_tracing,
                       denliq=denliq, denice=denice, fields=("tbtotal",))
        tb = res["tbtotal"].cpu().numpy().reshape(T, C, len(frqs), int(good.sum()))
        out[m][:, :, good, :] = tb.transpose(0, 2, 3, 1)
    return out


def derive_TBs4PyRTlib(ds, args=None):
    """Same name and contract as the reference function: adds TBs_PyRTlib_{R24,R17,R98,R20} with
    dims ('time','N_Channels','elevation','Crop') and the reference's attrs to `ds` and returns it.
    `ds` is an xarray.Dataset or an io_formats.SimpleDataset."""
    elev = np.asarray(ds["elevation"].values, dtype=np.float64)
    tbs = derive_tbs_arrays(_level_values(ds, "Level_z"), _level_values(ds, "Level_Pressure"),
                            _level_values(ds, "Level_Temperature"), _level_values(ds, "Level_RH"),
                            elevations=elev)
    for model, var in io_formats.OUTPUT_VARS:
        ds[var] = (io_formats.TB_DIMS, tbs[model])
        attrs = io_formats.tb_attrs(model)
        note = provenance(model)
        if note:     # the built-in line tables are not pinned to PyRTlib: say so in the product
            attrs["model_tables"] = note
        ds[var].attrs = attrs
    return ds


LEVEL_DIMS = ("N_Levels", "time", "Crop")   # preprocessing4all.py:1195-1203


def _level_values(ds, name):
    """`Level_*` as a (N_Levels, time, Crop) array.  The reference indexes BY NAME
    (`.isel(time=i).isel(Crop=j)`, PyRTlib_processing.py:109-114), so it does not depend on the stored
    axis order; a dataset that was transposed or merged time-major is put back in contract order here
    instead of being unpacked with levels and profiles swapped."""
    var = ds[name]
    dims = tuple(getattr(var, "dims", ()) or ())
    vals = np.asarray(var.values)
    if not dims or dims == LEVEL_DIMS:
        return vals
    if sorted(dims) != sorted(LEVEL_DIMS):
        raise ValueError("%s has dims %s, expected some order of %s" % (name, dims, LEVEL_DIMS))
    return np.transpose(vals, [dims.index(d) for d in LEVEL_DIMS])


def legacy_sonde_table(z_km, p, t, rh, z_km_crop, p_crop, t_crop, rh_crop):
    """The legacy runner's per-sonde product: zenith tbtotal for the 9 legacy models, cropped and
    uncropped profile (ground-first 1-D arrays) -> (cropped, uncropped) dicts for write_legacy_csv."""
    blocks = []
    for zz, pp, tt, rr in ((z_km_crop, p_crop, t_crop, rh_crop), (z_km, p, t, rh)):
        d = {}
        for m in io_formats.LEGACY_MODELS:
            res = simulate(np.asarray(zz)[None], np.asarray(pp)[None], np.asarray(tt)[None],
                           np.asarray(rr)[None], HATPRO_FRQ, [90.0], m, fields=("tbtotal",))
            d[m] = res["tbtotal"].cpu().numpy()[0, :, 0]
        blocks.append(d)
    return blocks[0], blocks[1]


def parse_arguments(argv=None):
    parser = argparse.ArgumentParser(description="Process radiosondes into PyRTlib-model TBs on the GPU")
    outpath = "~/PhD_data/TB_preproc_and_proc_results/"
    parser.add_argument("--input", "-i", type=str,
                        default=os.path.expanduser(outpath + "MWR_rs_FESSTVaLSoclesVital1_all_elevations.nc"),
                        help="NetCDF file with rs and MWR data")
    parser.add_argument("--output", "-o", type=str,
                        default=os.path.expanduser(outpath + "3_campaigns_PyRTlib_R24_processed_TBs_from_rs.nc"),
                        help="Where to save summarized NetCDF of Inputs and Output TBs")
    return parser.parse_args(argv)


def main(argv=None):
    args = parse_arguments(argv)
    ds = io_formats.open_dataset(args.input)
    ds = derive_TBs4PyRTlib(ds, args)
    try:
        ds.to_netcdf(args.output, format="NETCDF4_CLASSIC")
    except (TypeError, ValueError):
        ds.to_netcdf(args.output)


if __name__ == "__main__":
    main()
