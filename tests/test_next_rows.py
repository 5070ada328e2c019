"""N2 wire formats + the batched driver, N4 statistics.  CPU tests cover the
formats and the oracle ports; the GPU tests check the CUDA path against those ports."""
import os

import numpy as np
import pytest

from oracle import postproc


def _synth_level_dataset(T=3, C=2, L=24, seed=5):
    """A tiny stand-in for MWR_rs_..._all_elevations.nc: Level_* (N_Levels, time, Crop), top-first."""
    import mwr_fast_forward_operators_and_lbls_b200 as pkg
    d = pkg.synth_profiles(T * C, L, seed=seed)

    def lay(v, mul):   # (T*C, L) ground-first -> (L, T, C) top-first
        return np.ascontiguousarray((v * mul).T[::-1].reshape(L, T, C))
    return dict(Level_z=lay(d["z"], 1000.0), Level_Pressure=lay(d["p"], 1.0),
                Level_Temperature=lay(d["t"], 1.0), Level_RH=lay(d["rh"], 100.0))


def test_legacy_csv_roundtrip(tmp_path):
    from mwr_fast_forward_operators_and_lbls_b200 import io_formats as io
    rng = np.random.default_rng(0)
    crop = {m: rng.uniform(10, 290, 14) for m in io.LEGACY_MODELS}
    full = {m: rng.uniform(10, 290, 14) for m in io.LEGACY_MODELS}
    path = str(tmp_path / "sonde_out.txt")
    df = io.write_legacy_csv(path, crop, full)
    assert len(df) == 252 and list(df.columns) == ["tbtotal"]
    # the consumer's slicing (old_merge2nc.py:417-435)
    import pandas as pd
    prl = pd.read_csv(path)["tbtotal"].values
    assert np.allclose(prl[0:14], crop["R17"]) and np.allclose(prl[112:126], crop["R24"])
    assert np.allclose(prl[126:140], full["R17"]) and np.allclose(prl[238:252], full["R24"])
    c2, f2 = io.read_legacy_csv(path)
    assert all(np.allclose(c2[m], crop[m]) and np.allclose(f2[m], full[m]) for m in io.LEGACY_MODELS)


def test_netcdf_contract_roundtrip(tmp_path):
    from mwr_fast_forward_operators_and_lbls_b200 import io_formats as io
    lv = _synth_level_dataset()
    ds = io.SimpleDataset({k: (("N_Levels", "time", "Crop"), v) for k, v in lv.items()})
    ds["elevation"] = (("elevation",), np.array([90., 30., 4.2]))
    tb = np.random.default_rng(1).uniform(10, 290, (3, 14, 3, 2))
    for model, var in io.OUTPUT_VARS:
        ds[var] = (io.TB_DIMS, tb)
        ds[var].attrs = io.tb_attrs(model)
    path = str(tmp_path / "out.nc")
    ds.to_netcdf(path, format="NETCDF4_CLASSIC")
    back = io.open_dataset(path)
    for model, var in io.OUTPUT_VARS:
        assert tuple(back[var].dims) == ("time", "N_Channels", "elevation", "Crop")
        assert np.array_equal(np.asarray(back[var].values), tb)
        assert back[var].attrs["units"] == "K"
        assert back[var].attrs["long_name"] == "Brightness temperature modelled by %s" % model
        assert back[var].attrs["standard_name"] == "brightness_temperature"
    assert np.array_equal(np.asarray(back["Level_z"].values), lv["Level_z"])


def test_driver_reads_level_variables_by_dimension_name():
    """The reference indexes Level_* by name (`.isel(time=i).isel(Crop=j)`, PyRTlib_processing.py:109-114):
    a dataset stored time-major must be unpacked exactly like one in contract order, and an unknown
    dimension is an error instead of a silent level/profile swap."""
    from mwr_fast_forward_operators_and_lbls_b200 import driver, io_formats as io
    lv = _synth_level_dataset(T=3, C=2, L=24)
    ds = io.SimpleDataset({k: (("N_Levels", "time", "Crop"), v) for k, v in lv.items()})
    tm = io.SimpleDataset({k: (("time", "Crop", "N_Levels"), np.ascontiguousarray(v.transpose(1, 2, 0)))
                           for k, v in lv.items()})
    for k in lv:
        assert np.array_equal(driver._level_values(ds, k), lv[k])
        assert np.array_equal(driver._level_values(tm, k), lv[k])
    bad = io.SimpleDataset({"Level_z": (("N_Levels", "time", "member"), lv["Level_z"])})
    with pytest.raises(ValueError):
        driver._level_values(bad, "Level_z")


def test_builtin_model_tables_carry_their_provenance():
    """The built-in line tables are not pinned to PyRTlib output: the package says so once per model
    (warning) and the driver writes it into the output attributes."""
    from mwr_fast_forward_operators_and_lbls_b200 import models
    assert models.BUILTIN_TABLES_VERIFIED is False
    note = models.provenance("R24")
    assert "R24" in note and "NOT verified" in note


def test_stats_oracle_matches_definition():
    rng = np.random.default_rng(2)
    v = rng.normal(150, 30, (40, 14))
    r = v + rng.normal(0.3, 1.0, v.shape)
    v[3, 2] = np.nan
    r[7, 5] = np.nan
    sd, bias, rmse = postproc.stats_by_channel(v, r, 14)
    d = v[:, 0] - r[:, 0]
    assert np.isclose(bias[0], d.mean()) and np.isclose(rmse[0], np.sqrt((d ** 2).mean()))
    assert np.isclose(sd[0], d.std())
    # reference quirk: n_valid counts the model values only
    d5 = v[:, 5] - r[:, 5]
    assert np.isclose(bias[5], np.nansum(d5) / 40)


@pytest.mark.gpu
def test_stats_gpu_vs_oracle(built_lib):
    from mwr_fast_forward_operators_and_lbls_b200 import stats
    rng = np.random.default_rng(3)
    v = rng.normal(150, 30, (521, 14, 10))
    r = v + rng.normal(0.3, 1.0, v.shape)
    v[rng.uniform(size=v.shape) < 0.02] = np.nan
    r[rng.uniform(size=v.shape) < 0.02] = np.nan
    sd, bias, rmse = stats.stats_by_channel(v, r, 14)
    for k in range(10):
        o_sd, o_bias, o_rmse = postproc.stats_by_channel(v[:, :, k], r[:, :, k], 14)
        assert np.allclose(sd.cpu().numpy()[:, k], o_sd, rtol=1e-12, atol=0)
        assert np.allclose(bias.cpu().numpy()[:, k], o_bias, rtol=1e-10, atol=1e-13)
        assert np.allclose(rmse.cpu().numpy()[:, k], o_rmse, rtol=1e-12, atol=0)


@pytest.mark.gpu
def test_driver_output_attrs_carry_the_table_provenance(built_lib):
    from mwr_fast_forward_operators_and_lbls_b200 import driver, io_formats as io
    lv = _synth_level_dataset(T=2, C=2, L=24)
    ds = io.SimpleDataset({k: (("time", "Crop", "N_Levels"), np.ascontiguousarray(v.transpose(1, 2, 0)))
                           for k, v in lv.items()})               # time-major on purpose
    ds["elevation"] = (("elevation",), np.array([90., 30.]))
    ref = io.SimpleDataset({k: (("N_Levels", "time", "Crop"), v) for k, v in lv.items()})
    ref["elevation"] = (("elevation",), np.array([90., 30.]))
    ds, ref = driver.derive_TBs4PyRTlib(ds), driver.derive_TBs4PyRTlib(ref)
    for model, var in io.OUTPUT_VARS:
        assert "NOT verified" in ds[var].attrs["model_tables"] and ds[var].attrs["units"] == "K"
        assert np.array_equal(ds[var].values, ref[var].values)
        assert np.isfinite(ds[var].values).all()


@pytest.mark.gpu
def test_batched_driver_equals_reference_loop(built_lib, tmp_path):
    """derive_TBs4PyRTlib: four batched GPU calls == the reference's triple loop (oracle port),
    including the per-profile NaN gate and the output variable contract."""
    from mwr_fast_forward_operators_and_lbls_b200 import driver, io_formats as io
    import mwr_fast_forward_operators_and_lbls_b200 as pkg
    lv = _synth_level_dataset(T=2, C=2, L=20)
    lv["Level_Temperature"][5, 1, 0] = np.nan           # one NaN profile (time 1, Crop 0)
    elev = np.array([90., 19.2, 4.2])
    ref = postproc.derive_TBs4PyRTlib_loop(lv["Level_z"], lv["Level_Pressure"], lv["Level_Temperature"],
                                           lv["Level_RH"], elev, pkg.HATPRO_FRQ)
    ds = io.SimpleDataset({k: (("N_Levels", "time", "Crop"), v) for k, v in lv.items()})
    ds["elevation"] = (("elevation",), elev)
    ds = driver.derive_TBs4PyRTlib(ds, None)
    for model, var in io.OUTPUT_VARS:
        got = ds[var].values
        assert got.shape == (2, 14, 3, 2) and ds[var].dims == io.TB_DIMS
        assert np.isnan(got[1, :, :, 0]).all() and np.isnan(ref[model][1, :, :, 0]).all()
        ok = ~np.isnan(ref[model])
        assert np.array_equal(ok, ~np.isnan(got))
        assert np.max(np.abs(got[ok] - ref[model][ok])) < 1e-3
        assert ds[var].attrs["units"] == "K"
    # CLI round trip (-i/-o like the reference)
    src, dst = str(tmp_path / "in.nc"), str(tmp_path / "out.nc")
    ds_in = io.SimpleDataset({k: (("N_Levels", "time", "Crop"), v) for k, v in lv.items()})
    ds_in["elevation"] = (("elevation",), elev)
    ds_in.to_netcdf(src)
    driver.main(["-i", src, "-o", dst])
    back = io.open_dataset(dst)
    assert np.allclose(np.asarray(back["TBs_PyRTlib_R24"].values), ds["TBs_PyRTlib_R24"].values,
                       equal_nan=True)


@pytest.mark.gpu
def test_legacy_sonde_table(built_lib, tmp_path):
    from mwr_fast_forward_operators_and_lbls_b200 import driver, io_formats as io
    import mwr_fast_forward_operators_and_lbls_b200 as pkg
    from oracle import pyrtlib_vec as pv
    d = pkg.synth_profiles(1, 30, seed=9)
    z, p, t, rh = (d[k][0] for k in ("z", "p", "t", "rh"))
    crop, full = driver.legacy_sonde_table(z, p, t, rh, z[2:], p[2:], t[2:], rh[2:])
    path = str(tmp_path / "x_out.txt")
    io.write_legacy_csv(path, crop, full)
    c2, f2 = io.read_legacy_csv(path)
    for m in ("R17", "R98", "R24"):
        ref = pv.simulate(z[None], p[None], t[None], rh[None], pkg.HATPRO_FRQ, [90.0], m)["tbtotal"][0, :, 0]
        assert np.max(np.abs(f2[m] - ref)) < 1e-3


@pytest.mark.gpu
def test_cloudy_wiring_from_level_liquid(built_lib):
    """N3: Level_Liquid / Level_Ice [kg/kg] -> g m^-3 -> cloudy path; against the oracle."""
    from mwr_fast_forward_operators_and_lbls_b200 import driver
    import mwr_fast_forward_operators_and_lbls_b200 as pkg
    from oracle import pyrtlib_vec as pv
    T, C, L = 3, 2, 40
    d = pkg.synth_profiles(T * C, L, seed=12, clouds=True)
    rho = d["p"] * 100.0 / 287.06 / d["t"]                     # kg m^-3
    liq_kgkg = d["denliq"] / 1000.0 / rho                      # what preprocessing stores

    def lay(v, mul=1.0):
        return np.ascontiguousarray((v * mul).T[::-1].reshape(L, T, C))
    elev = np.array([90.0, 5.4])
    got = driver.derive_tbs_arrays(lay(d["z"], 1000.0), lay(d["p"]), lay(d["t"]), lay(d["rh"], 100.0),
                                   elevations=elev, models=("R24",), level_liquid=lay(liq_kgkg))["R24"]
    denliq = driver.water_content_to_density(liq_kgkg, d["p"], d["t"])
    assert np.allclose(denliq, d["denliq"], rtol=1e-12, atol=1e-15)
    ref = pv.simulate(d["z"] * 1000.0 / 1000.0, d["p"], d["t"], d["rh"] * 100.0 / 100.0, pkg.HATPRO_FRQ,
                      elev, "R24", denliq=denliq, denice=np.zeros_like(denliq))["tbtotal"]
    ref = ref.reshape(T, C, 14, 2).transpose(0, 2, 3, 1)
    assert (d["denliq"] > 0).any()
    assert np.max(np.abs(got - ref)) < 1e-3


# ---- N1: profile pre-processing (thinning, running means, climatology top-up, lowest-180) ----------
GOLD_PRE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "preproc_reference.npz")
_LEVEL_KEYS = ("p", "t", "ppmv", "h2o_gkg", "z", "rh")


def _pre_cases():
    g = np.load(GOLD_PRE)
    clim = {k: g["clim_" + k] for k in ("z_km", "p", "t", "gkg", "rh")}
    return g, clim, list(zip(g["keys"], g["rejected"]))


def test_preproc_oracle_equals_reference_output():
    """oracle/preproc.py == the reference's own functions (outputs recorded by
    tests/golden/make_preproc_golden.py from /root/reference's preprocessing4all.py): bit-exact."""
    from oracle import preproc as op
    g, clim, cases = _pre_cases()
    seen_10hpa = False
    for key, rej in cases:
        h, t, p, rh = g[key.rstrip("v") + "_raw"]
        inds = op.thinning_indices(h, p)
        r = op.assemble_levels(h, t, p, rh, inds, clim, z_offset=20.0 if key.endswith("v") else 0.0)
        assert r["ok"] == (not rej)
        if rej:
            assert all(np.isnan(r[k]).all() for k in _LEVEL_KEYS)
            continue
        zt, tt, pt, rt = g[key + "_thinned"]
        assert np.array_equal(h[inds] - (20.0 if key.endswith("v") else 0.0), zt)
        assert np.array_equal(op.running_mean_from_arrays(inds, t), tt, equal_nan=True)
        got = np.stack([r[k] for k in _LEVEL_KEYS])
        assert np.array_equal(got, g[key + "_levels"], equal_nan=True), key
        # lowest2tenhPa fired if the 7 topmost kept levels are not a contiguous run of the column
        seen_10hpa = seen_10hpa or got[0, 0] < 10.0
    assert seen_10hpa


def test_preproc_host_logic_matches_oracle():
    from oracle import preproc as op
    from mwr_fast_forward_operators_and_lbls_b200 import preproc as pp
    g, clim, cases = _pre_cases()
    for key, _ in cases:
        h, t, p, rh = g[key.rstrip("v") + "_raw"]
        assert op.detect_crop(h) == pp.detect_crop(h)
        a, b = op.thinning_indices(h, p), pp.thinning_indices(h, p)
        assert np.array_equal(a, b)
    short = g["c0_raw"][:, :250]          # fewer than 300 usable samples -> rejected (:246)
    assert pp.thinning_indices(short[0], short[2]) is None and op.thinning_indices(short[0], short[2]) is None
    low = g["c0_raw"][:, :900]            # 10 km not reached (:251)
    assert pp.thinning_indices(low[0], low[2]) is None
    # the climatology the product ships is the one the goldens were generated with
    c = pp.midlatitude_summer()
    assert all(np.array_equal(c[k], clim[k]) for k in clim)


@pytest.mark.gpu
def test_preproc_gpu_vs_reference_output(built_lib):
    """CUDA thinning/assembly vs the reference's recorded outputs and vs the oracle."""
    from mwr_fast_forward_operators_and_lbls_b200 import preproc as pp
    g, clim, cases = _pre_cases()
    N = max(g[k + "_raw"].shape[1] for k, _ in cases if not k.endswith("v"))

    def pad(a):
        return np.concatenate([a, np.full(N - a.size, np.nan)])
    raws = [g[k.rstrip("v") + "_raw"] for k, _ in cases]
    H, T, P, R = (np.stack([pad(r[i]) for r in raws]) for i in range(4))
    vital = np.array([k.endswith("v") for k, _ in cases])
    out = pp.preprocess_sondes(H, T, P, R, vital=vital)
    ok = out["ok"].cpu().numpy()
    names = ("Level_Pressure", "Level_Temperature", "Level_ppmvs", "Level_H2O", "Level_z", "Level_RH")
    for s, (key, rej) in enumerate(cases):
        assert ok[s] == (0 if rej else 1)
        got = np.stack([out[n][s].cpu().numpy() for n in names])
        if rej:
            assert np.isnan(got).all()
            continue
        ref = g[key + "_levels"]
        assert np.array_equal(np.isnan(got), np.isnan(ref))
        m = ~np.isnan(ref)
        rel = np.abs(got[m] - ref[m]) / np.maximum(np.abs(ref[m]), 1e-300)
        assert rel.max() < 1e-12, (key, rel.max())
        n = int(out["thinned"]["n"][s])
        th = np.stack([out["thinned"][k][s, :n].cpu().numpy() for k in ("z", "t", "p", "rh")])
        rt = g[key + "_thinned"]
        if np.all(th[3] <= 1.5):          # the reference returns RH already scaled to % (:299-300);
            th[3] = th[3] * 100           # the device applies that scaling in the assembly kernel
        assert np.array_equal(th[0], rt[0])
        mm = ~np.isnan(rt)
        assert np.array_equal(np.isnan(th), np.isnan(rt))
        assert (np.abs(th[mm] - rt[mm]) <= 1e-13 * np.abs(rt[mm])).all()


@pytest.mark.gpu
def test_preproc_feeds_the_forward_model(built_lib):
    """N1 -> hot path: Level_* from the device pre-processing go straight into the batched driver
    (top-first, %, m) and give the TBs the oracle gets from the reference-recorded levels."""
    from mwr_fast_forward_operators_and_lbls_b200 import preproc as pp, driver
    import mwr_fast_forward_operators_and_lbls_b200 as pkg
    from oracle import pyrtlib_vec as pv
    g, clim, cases = _pre_cases()
    h, t, p, rh = g["c0_raw"]
    out = pp.preprocess_sondes(h[None], t[None], p[None], rh[None])

    def lay(name):   # (1,180) top-first -> (180, time=1, Crop=1)
        return out[name][0].cpu().numpy().reshape(180, 1, 1)
    elev = np.array([90.0, 30.0, 4.2])
    got = driver.derive_tbs_arrays(lay("Level_z"), lay("Level_Pressure"), lay("Level_Temperature"),
                                   lay("Level_RH"), elevations=elev, models=("R24",))["R24"]
    lv = g["c0_levels"]
    ref = pv.simulate(lv[4][::-1][None] / 1000.0, lv[0][::-1][None], lv[1][::-1][None],
                      lv[5][::-1][None] / 100.0, pkg.HATPRO_FRQ, elev, "R24")["tbtotal"]
    assert np.max(np.abs(got[0, :, :, 0] - ref[0])) < 1e-3


# ---- N3: batched cloud detection + adiabatic water content (derive_cloud_water.py) ------------------
GOLD_CLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cloud_reference.npz")


def test_cloud_oracle_equals_reference_output():
    """oracle/cloud.py == the reference's own derive_cloud_features (outputs recorded by
    tests/golden/make_cloud_golden.py from /root/reference's derive_cloud_water.py): bit-exact."""
    from oracle import cloud
    g = np.load(GOLD_CLD)
    n_liq = n_ice = 0
    for c in range(len(g["p"])):
        o = cloud.derive_cloud_features(g["p"][c], g["t"][c], g["z"][c], g["rh"][c])
        for a, k in ((o[0], "lwc_m3"), (o[1], "lwc_kg"), (o[3], "iwc_m3"), (o[4], "iwc_kg")):
            assert np.array_equal(a, g[k][c], equal_nan=True), (c, k)
        assert o[2] == g["lwp"][c] and o[5] == g["iwp"][c]
        n_liq += bool((o[0] > 0).any())
        n_ice += bool((o[3] > 0).any())
    assert n_liq > 50 and n_ice > 50          # the goldens exercise both phases


@pytest.mark.gpu
def test_cloud_features_gpu_vs_reference_output(built_lib):
    from mwr_fast_forward_operators_and_lbls_b200 import preproc as pp
    g = np.load(GOLD_CLD)
    out = pp.derive_cloud_features(g["p"], g["t"], g["z"], g["rh"])
    for k, gk in (("lwc_kg_m3", "lwc_m3"), ("lwc_kg_kg", "lwc_kg"), ("iwc_kg_m3", "iwc_m3"),
                  ("iwc_kg_kg", "iwc_kg"), ("lwp", "lwp"), ("iwp", "iwp")):
        got, ref = out[k].cpu().numpy(), g[gk]
        assert np.array_equal(got == 0.0, ref == 0.0), k          # same cloudy levels
        nz = ref != 0.0
        assert np.max(np.abs(got[nz] - ref[nz]) / np.abs(ref[nz])) < 1e-12, k


@pytest.mark.gpu
def test_cloud_features_feed_the_cloudy_forward_model(built_lib):
    """N3 end to end: RH profile -> device cloud detection -> Level_Liquid -> cloudy TbCloudRTE path,
    against the oracle chain (oracle cloud detection -> oracle forward model)."""
    from mwr_fast_forward_operators_and_lbls_b200 import preproc as pp, driver
    import mwr_fast_forward_operators_and_lbls_b200 as pkg
    from oracle import cloud, pyrtlib_vec as pv
    g = np.load(GOLD_CLD)
    sel = [c for c in range(len(g["p"])) if (g["lwc_m3"][c] > 0).any()][:6]
    p, t, z, rh = (g[k][sel] for k in ("p", "t", "z", "rh"))
    out = pp.derive_cloud_features(p, t, z, rh)
    T, L = len(sel), p.shape[1]

    def lay(a):   # (T,L) top-first -> (L, time=T, Crop=1)
        return np.ascontiguousarray(a.T.reshape(L, T, 1))
    elev = np.array([90.0, 19.2])
    got = driver.derive_tbs_arrays(lay(z), lay(p), lay(t), lay(np.minimum(rh, 100.0)), elevations=elev,
                                   models=("R24",), level_liquid=lay(out["lwc_kg_kg"].cpu().numpy()),
                                   level_ice=lay(out["iwc_kg_kg"].cpu().numpy()))["R24"]
    lw = np.array([cloud.derive_cloud_features(p[i], t[i], z[i], rh[i])[1] for i in range(T)])
    iw = np.array([cloud.derive_cloud_features(p[i], t[i], z[i], rh[i])[4] for i in range(T)])
    gf = lambda a: np.ascontiguousarray(a[:, ::-1])           # ground-first
    denl = driver.water_content_to_density(gf(lw), gf(p), gf(t))
    deni = driver.water_content_to_density(gf(iw), gf(p), gf(t))
    ref = pv.simulate(gf(z) / 1000.0, gf(p), gf(t), gf(np.minimum(rh, 100.0)) / 100.0, pkg.HATPRO_FRQ, elev,
                      "R24", denliq=denl, denice=deni)["tbtotal"]
    assert np.max(np.abs(got[:, :, :, 0] - ref)) < 1e-3
    clear = pv.simulate(gf(z) / 1000.0, gf(p), gf(t), gf(np.minimum(rh, 100.0)) / 100.0, pkg.HATPRO_FRQ, elev,
                        "R24")["tbtotal"]
    assert np.max(np.abs(ref - clear)) > 0.5      # the detected clouds matter at 31.4 GHz


def test_stats_oracle_equals_reference_output():
    """oracle stats_by_channel == the reference's own function (tests/golden/make_stats_golden.py),
    NaN gaps in values and references included: bit-exact."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "stats_reference.npz"))
    for c in range(int(g["n"])):
        v, r = g["v%d" % c], g["r%d" % c]
        sd, bias, rmse = postproc.stats_by_channel(v, r, v.shape[1])
        assert np.array_equal(sd, g["sd%d" % c], equal_nan=True)
        assert np.array_equal(bias, g["bias%d" % c], equal_nan=True)
        assert np.array_equal(rmse, g["rmse%d" % c], equal_nan=True)


@pytest.mark.gpu
def test_stats_gpu_vs_reference_output(built_lib):
    from mwr_fast_forward_operators_and_lbls_b200 import stats
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "stats_reference.npz"))
    for c in range(int(g["n"])):
        v, r = g["v%d" % c], g["r%d" % c]
        sd, bias, rmse = stats.stats_by_channel(v, r, v.shape[1])
        assert np.allclose(sd.cpu().numpy().ravel(), g["sd%d" % c], rtol=1e-12, atol=0, equal_nan=True)
        assert np.allclose(bias.cpu().numpy().ravel(), g["bias%d" % c], rtol=1e-10, atol=1e-13, equal_nan=True)
        assert np.allclose(rmse.cpu().numpy().ravel(), g["rmse%d" % c], rtol=1e-12, atol=0, equal_nan=True)
