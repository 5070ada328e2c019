"""GPU parity tests: the CUDA path (through the C ABI of include/mwrt.h) against the CPU oracle on the
same seeded inputs.  Tolerances are the ones BASELINE.json:north_star states:
  fp64 absorption coefficients and optical depths within 1e-9 relative, TB within 1e-3 K.
The oracle is the in-repo restatement (parity with real PyRTlib is unpinned).
"""
import copy
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

REL = 1e-9      # absorption / optical depth, relative (north_star)
TB_ABS = 1e-3   # brightness temperature, K (north_star)
GOLD = os.path.join(os.path.dirname(__file__), "golden", "config1_us_standard.json")


@pytest.fixture(scope="module")
def pkg(built_lib):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import mwr_fast_forward_operators_and_lbls_b200 as m
    return m


@pytest.fixture(scope="module")
def ora():
    from oracle import atmospheres, pyrtlib_scalar, pyrtlib_vec
    class O:  # noqa: E306
        atm, ps, pv = atmospheres, pyrtlib_scalar, pyrtlib_vec
    return O


def rel_err(a, b, floor=0.0):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor + 1e-300)))


def cpu(x):
    return x.detach().cpu().numpy()


ALL_MODELS = ["R98", "R03", "R16", "R17", "R19", "R19SD", "R20", "R20SD", "R21SD", "R22SD", "R23",
              "R23SD", "R24"]


def test_native_library_loaded(pkg):
    """The extension in use is the in-tree libmwrt.so (no eager/CPU fallback exists)."""
    from mwr_fast_forward_operators_and_lbls_b200 import _lib
    maps = open("/proc/self/maps").read()
    _lib.load()
    assert "libmwrt.so" in open("/proc/self/maps").read() or "libmwrt.so" in maps


def test_vapor(pkg, ora):
    d = pkg.synth_profiles(64, 100, seed=3)
    e, rho = pkg.vapor(d["t"], d["rh"])
    e0, rho0 = ora.pv.vapor(d["t"], d["rh"])
    assert rel_err(cpu(e), e0) < 1e-13
    assert rel_err(cpu(rho), rho0) < 1e-13


@pytest.mark.parametrize("mdl", ALL_MODELS)
def test_absorption_hatpro_channels(pkg, ora, mdl):
    """awet/adry at the 14 HATPRO channels, every model string, 1e-9 relative."""
    d = pkg.synth_profiles(48, 100, seed=11)
    e, _ = ora.pv.vapor(d["t"], d["rh"])
    g = pkg.absorption(mdl, d["p"], d["t"], e, pkg.HATPRO_FRQ)
    awet, adry, _, _ = ora.pv.absorption(mdl, d["p"], d["t"], e, pkg.HATPRO_FRQ)
    assert rel_err(cpu(g["awet"]), awet) < REL
    assert rel_err(cpu(g["adry"]), adry) < REL
    assert int(g["status"].sum()) == 0


@pytest.mark.parametrize("mdl", ["R98", "R17", "R20SD", "R24"])
def test_absorption_wide_spectrum(pkg, ora, mdl):
    """1-1000 GHz sweep: exercises the 750 GHz cutoffs, line centres, the speed-dependent 22/183 GHz
    and 118 GHz shapes, the tail frequency block (F=37 is not a multiple of 16)."""
    d = pkg.synth_profiles(6, 60, seed=5)
    e, _ = ora.pv.vapor(d["t"], d["rh"])
    frq = np.concatenate([np.linspace(1.0, 1000.0, 25), [22.235, 118.7503, 183.31, 60.3061, 556.936,
                                                        752.0332, 916.1712, 50.0, 89.0, 150.0, 325.1529,
                                                        2.0]])
    assert len(frq) == 37
    g = pkg.absorption(mdl, d["p"], d["t"], e, frq)
    awet, adry, _, _ = ora.pv.absorption(mdl, d["p"], d["t"], e, frq)
    assert rel_err(cpu(g["awet"]), awet) < REL
    assert rel_err(cpu(g["adry"]), adry) < REL


def test_absorption_custom_model_all_terms(pkg, ora):
    """Every table column non-zero (perturbed R22SD): shift terms, aair/aself, d2, g/dnu ..."""
    rng = np.random.default_rng(0)
    desc = copy.deepcopy(pkg.load_descriptor("R22SD"))
    desc["name"] = "perturbed"
    n = len(desc["h2o"]["fl"])
    for col, scale in (("sr", 1e-3), ("xh", 1.0), ("xhs", 1.0), ("aair", 0.5), ("aself", 0.5),
                       ("d2", 2e-5), ("d2s", 2e-4), ("sh", 1e-4), ("shs", 1e-3)):
        desc["h2o"][col] = (np.array(desc["h2o"][col]) + scale * rng.uniform(0.1, 1.0, n)).tolist()
    desc["h2o"]["w2"] = (np.array(desc["h2o"]["w2"]) + 2e-4 * (rng.uniform(0, 1, n) < 0.4)).tolist()
    m = len(desc["o2"]["f"])
    for col in ("g0", "g1", "dnu0", "dnu1"):
        desc["o2"][col] = (np.array(desc["o2"][col]) + 1e-3 * rng.uniform(-1, 1, m)).tolist()
    d = pkg.synth_profiles(8, 80, seed=21)
    e, _ = ora.pv.vapor(d["t"], d["rh"])
    frq = np.concatenate([pkg.HATPRO_FRQ, [118.0, 118.75, 183.0, 183.31, 321.0, 380.2]])
    g = pkg.absorption(desc, d["p"], d["t"], e, frq)
    awet, adry, _, _ = ora.pv.absorption(desc, d["p"], d["t"], e, frq)
    assert rel_err(cpu(g["awet"]), awet) < REL
    assert rel_err(cpu(g["adry"]), adry) < REL


@pytest.mark.parametrize("mdl", ["R98", "R24"])
def test_cloud_absorption(pkg, ora, mdl):
    d = pkg.synth_profiles(64, 100, seed=9, clouds=True)
    d["denice"][:, 40:43] = 0.03
    e, _ = ora.pv.vapor(d["t"], d["rh"])
    frq = np.array([22.24, 31.4, 58.0, 90.0, 150.0])
    g = pkg.absorption(mdl, d["p"], d["t"], e, frq, d["denliq"], d["denice"])
    _, _, aliq, aice = ora.pv.absorption(mdl, d["p"], d["t"], e, frq, d["denliq"], d["denice"])
    assert (aliq > 0).any()
    assert rel_err(cpu(g["aliq"]), aliq) < REL
    assert rel_err(cpu(g["aice"]), aice) < REL


@pytest.mark.parametrize("rt", [False, True])
def test_slant_path(pkg, ora, rt):
    d = pkg.synth_profiles(32, 100, seed=13)
    e, _ = ora.pv.vapor(d["t"], d["rh"])
    ang = np.concatenate([pkg.HATPRO_ELEV, [89.5, 1.0, 45.0]])
    ds, st = pkg.slant_path(d["z"], d["p"], d["t"], e, ang, rt)
    ds0 = ora.pv.slant_path(d["z"], d["p"], d["t"], e, ang, rt)
    assert rel_err(cpu(ds)[:, :, 1:], ds0[:, :, 1:]) < REL
    assert np.all(cpu(ds)[:, :, 0] == 0.0)


def test_slant_path_ducting_long_profiles_and_many_angles(pkg, ora):
    """Warp-cooperative slant path: (a) an elevated duct (humidity collapse 0.4 km up) seen at 0.5 deg --
    the path ends there, the rest of ds is 0 and the profile carries MWRT_ST_DUCTING, exactly where the
    oracle's serial loop stops; (b) L = 150 (five rounds of 32 levels, ragged last round) and 20 angles
    (more angles than warps)."""
    d = pkg.synth_profiles(6, 150, seed=14)
    d["rh"][2, :10] = 1.0
    d["rh"][2, 10:] = 0.01
    d["t"][2, 10:] += 6.0                                   # warm dry air above a moist layer: duct
    e, _ = ora.pv.vapor(d["t"], d["rh"])
    ang = np.concatenate([pkg.HATPRO_ELEV, [89.5, 1.0, 45.0, 0.5, 0.2, 0.1, 60.0, 75.0, 3.0, 2.0]])
    ds, st = pkg.slant_path(d["z"], d["p"], d["t"], e, ang, True)
    ds0 = ora.pv.slant_path(d["z"], d["p"], d["t"], e, ang, True)
    got = cpu(ds)
    assert ((got == 0) == (ds0 == 0)).all()
    m = ds0 != 0
    assert np.max(np.abs(got[m] - ds0[m]) / ds0[m]) < REL
    ducted = [(ds0[i][:, 1:] == 0).any() for i in range(6)]
    assert any(ducted)
    assert [bool(v & 4) for v in cpu(st)] == ducted


@pytest.mark.parametrize("from_sat", [False, True])
def test_rte_stage(pkg, ora, from_sat):
    """exponential_integration + planck + bright from the ORACLE's absorption: isolates kernel 3."""
    import torch
    d = pkg.synth_profiles(24, 100, seed=17)
    e, _ = ora.pv.vapor(d["t"], d["rh"])
    awet, adry, _, _ = ora.pv.absorption("R19", d["p"], d["t"], e, pkg.HATPRO_FRQ)
    ds = ora.pv.slant_path(d["z"], d["p"], d["t"], e, pkg.HATPRO_ELEV, False)
    g = pkg.rte(d["t"], torch.from_numpy(awet).cuda(), torch.from_numpy(adry).cuda(),
                torch.from_numpy(ds).cuda(), pkg.HATPRO_FRQ, from_sat=from_sat, tau_layer_angle=3)
    ref = ora.pv.simulate(d["z"], d["p"], d["t"], d["rh"], pkg.HATPRO_FRQ, pkg.HATPRO_ELEV, "R19",
                          from_sat=from_sat, return_layers=True)
    for k in ("tauwet", "taudry"):
        assert rel_err(cpu(g[k]), ref[k]) < REL, k
    assert rel_err(cpu(g["tau_layer"])[:, :, 1:], ref["tau_layer"][:, :, 3, 1:]) < REL
    for k in ("tbtotal", "tbatm", "tmr"):
        assert np.max(np.abs(cpu(g[k]) - ref[k])) < TB_ABS, k


@pytest.mark.parametrize("mdl", ["R98", "R17", "R20", "R24", "R22SD"])
@pytest.mark.parametrize("rt", [False, True])
def test_end_to_end_config2_shape(pkg, ora, mdl, rt):
    """BASELINE.json config 2 at oracle-sized P: 100 levels x 14 channels x 10 elevations."""
    d = pkg.synth_profiles(96, 100, seed=1234)
    out = pkg.simulate(d["z"], d["p"], d["t"], d["rh"], pkg.HATPRO_FRQ, pkg.HATPRO_ELEV, mdl,
                       ray_tracing=rt)
    ref = ora.pv.simulate(d["z"], d["p"], d["t"], d["rh"], pkg.HATPRO_FRQ, pkg.HATPRO_ELEV, mdl,
                          ray_tracing=rt)
    for k in ("tauwet", "taudry"):
        assert rel_err(cpu(out[k]), ref[k]) < REL, k
    for k in ("tbtotal", "tbatm", "tmr"):
        err = np.max(np.abs(cpu(out[k]) - ref[k]))
        assert err < TB_ABS, (k, err)
    assert int(out["status"].sum()) == 0


def test_golden_config1(pkg):
    """BASELINE.json config 1 against the committed vectors (generated by tests/golden/make_golden.py)."""
    g = json.load(open(GOLD))
    z, p, t, rh = (np.array(g[k]) for k in ("z", "p", "t", "rh"))
    frq, ang = np.array(g["frq"]), np.array(g["angles"])
    for case in g["cases"]:
        out = pkg.simulate(z[None], p[None], t[None], rh[None], frq, ang, case["model"],
                           ray_tracing=case["ray_tracing"])
        for k in ("tbtotal", "tbatm", "tmr"):
            a = np.array(case[k]).reshape(len(ang), len(frq)).T
            assert np.max(np.abs(cpu(out[k])[0] - a)) < TB_ABS, (case["model"], k)
        for k in ("tauwet", "taudry"):
            a = np.array(case[k]).reshape(len(ang), len(frq)).T
            assert rel_err(cpu(out[k])[0], a) < REL, (case["model"], k)


def test_tbcloudrte_class_matches_scalar_oracle(pkg, ora):
    """The reference's exact call pattern (PyRTlib_processing.py:123-127): top-first arrays flipped
    with [::-1], ONE angle per object, tbtotal column of the DataFrame."""
    z, p, t, rh = ora.atm.us_standard()
    z_in, p_in, t_in, rh_in = z[::-1].copy(), p[::-1].copy(), t[::-1].copy(), rh[::-1].copy()
    keep = [v.copy() for v in (z_in, p_in, t_in, rh_in)]
    frqs = pkg.HATPRO_FRQ
    for elevation in (90.0, 5.4):
        ang = np.array([elevation])
        for mdl in ("R20", "R24", "R17", "R98"):
            rte = pkg.TbCloudRTE(z_in[::-1], p_in[::-1], t_in[::-1], rh_in[::-1], frqs, ang)
            rte.init_absmdl(mdl)
            rte.satellite = False
            df = rte.execute()
            o = ora.ps.TbCloudRTE(z_in[::-1], p_in[::-1], t_in[::-1], rh_in[::-1], frqs, ang)
            o.init_absmdl(mdl)
            o.satellite = False
            ref = o.execute()
            assert list(df.columns[:8]) == ["tbtotal", "tbatm", "tmr", "tmrcld", "tauwet", "taudry",
                                            "tauliq", "tauice"]
            assert np.max(np.abs(df["tbtotal"].values - ref["tbtotal"])) < TB_ABS
            assert rel_err(df["tauwet"].values, ref["tauwet"]) < REL
            assert rte.cloudy is False
    for a, b in zip(keep, (z_in, p_in, t_in, rh_in)):
        assert np.array_equal(a, b)  # inputs never modified
    # multi-angle object: DataFrame rows are angle-major, frequency fastest
    rte = pkg.TbCloudRTE(z, p, t, rh, frqs, pkg.HATPRO_ELEV, absmdl="R24", from_sat=False)
    df = rte.execute()
    assert len(df) == 140 and df["angle"].values[13] == 90.0 and df["angle"].values[14] == 30.0
    assert df["frq"].values[15] == frqs[1]
    with pytest.raises(ValueError):
        pkg.TbCloudRTE(z, p, t, rh, frqs, [90.0]).init_absmdl("R99")
    with pytest.raises(ValueError):
        pkg.TbCloudRTE(z, p, t, rh, frqs, [90.0]).execute()


def test_cloudy_end_to_end_and_class(pkg, ora):
    d = pkg.synth_profiles(40, 100, seed=77, clouds=True)
    out = pkg.simulate(d["z"], d["p"], d["t"], d["rh"], pkg.HATPRO_FRQ, pkg.HATPRO_ELEV, "R24",
                       denliq=d["denliq"], denice=d["denice"])
    ref = ora.pv.simulate(d["z"], d["p"], d["t"], d["rh"], pkg.HATPRO_FRQ, pkg.HATPRO_ELEV, "R24",
                          denliq=d["denliq"], denice=d["denice"])
    assert (ref["tauliq"] > 0).any()
    assert rel_err(cpu(out["tauliq"]), ref["tauliq"], floor=1e-30) < REL
    assert np.max(np.abs(cpu(out["tbtotal"]) - ref["tbtotal"])) < TB_ABS
    i = int(np.argmax(d["denliq"].max(axis=1)))
    rte = pkg.TbCloudRTE(d["z"][i], d["p"][i], d["t"][i], d["rh"][i], pkg.HATPRO_FRQ, [90.0, 30.0],
                         from_sat=False)
    rte.init_absmdl("R24")
    rte.init_cloudy(np.array([[1.0], [2.0]]), d["denice"][i], d["denliq"][i])
    df = rte.execute()
    assert np.max(np.abs(df["tbtotal"].values.reshape(2, 14).T - ref["tbtotal"][i][:, :2])) < TB_ABS
    assert (df["tauliq"].values > 0).all()


def test_nan_gate_per_profile(pkg, ora):
    """NaN in -> NaN out for that profile only (PyRTlib_processing.py:71-79,117-119,153-154)."""
    d = pkg.synth_profiles(16, 50, seed=2)
    good = pkg.simulate(d["z"], d["p"], d["t"], d["rh"], pkg.HATPRO_FRQ, pkg.HATPRO_ELEV[:3], "R98")
    for j, key in enumerate(("z", "p", "t", "rh")):
        d[key][3 + j, 10 + j] = np.nan
    out = pkg.simulate(d["z"], d["p"], d["t"], d["rh"], pkg.HATPRO_FRQ, pkg.HATPRO_ELEV[:3], "R98")
    st = cpu(out["status"])
    for i in range(16):
        bad = 3 <= i <= 6
        for k in pkg.FIELDS:
            if k in ("tauliq", "tauice") and not bad:
                continue
            v = cpu(out[k])[i]
            assert np.isnan(v).all() if bad else np.array_equal(v, cpu(good[k])[i]), (i, k)
        assert bool(st[i] & 1) == bad


def test_negative_absorption_flag_and_value_error(pkg):
    import torch
    d = pkg.synth_profiles(2, 20, seed=4)
    awet = torch.full((2, 3, 20), 1e-3, dtype=torch.float64, device="cuda")
    adry = awet.clone()
    adry[1, 2, 5] = -1.0
    ds = torch.ones((2, 1, 20), dtype=torch.float64, device="cuda")
    g = pkg.rte(d["t"], awet, adry, ds, [22.24, 31.4, 58.0])
    st = cpu(g["status"])
    assert st[0] == 0 and st[1] & 2
    assert np.isnan(cpu(g["tbtotal"])[1, 2]).all() and np.isfinite(cpu(g["tbtotal"])[0]).all()


def test_bandpass_fused_plan(pkg, ora):
    """BASELINE.json config 3 shape: sub-frequencies x top-hat average -> channels."""
    d = pkg.synth_profiles(12, 100, seed=8, clouds=True)
    nsub = 5
    frq = pkg.band_frequencies(nsub)
    mono = ora.pv.simulate(d["z"], d["p"], d["t"], d["rh"], frq, pkg.HATPRO_ELEV, "R24",
                           denliq=d["denliq"], denice=d["denice"])
    ref = mono["tbtotal"].reshape(12, 14, nsub, 10).mean(axis=2)
    out = pkg.simulate(d["z"], d["p"], d["t"], d["rh"], frq, pkg.HATPRO_ELEV, "R24", denliq=d["denliq"],
                       denice=d["denice"], bandpass=(14, nsub, None))
    assert tuple(out["tbtotal"].shape) == (12, 14, 10)
    assert np.max(np.abs(cpu(out["tbtotal"]) - ref)) < TB_ABS
    w = np.tile(np.array([0.1, 0.2, 0.4, 0.2, 0.1]), (14, 1))
    out = pkg.simulate(d["z"], d["p"], d["t"], d["rh"], frq, pkg.HATPRO_ELEV, "R24", denliq=d["denliq"],
                       denice=d["denice"], bandpass=(14, nsub, w))
    ref = (mono["tbtotal"].reshape(12, 14, nsub, 10) * w[None, :, :, None]).sum(axis=2)
    assert np.max(np.abs(cpu(out["tbtotal"]) - ref)) < TB_ABS


@pytest.mark.parametrize("nsub,weighted", [(5, False), (50, False), (50, True), (7, True), (400, False)])
def test_bandpass_fused_equals_unfused_bitwise(pkg, nsub, weighted):
    """The band-pass reduced inside kernel 3 == monochromatic fields + mwrt_bandpass_f64, bit for bit,
    for every field (same fma chain in the same order); nsub = 400 > 384 threads takes the plan's
    unfused fallback, nsub = 50 is BASELINE.json config 3 (7 channels per CTA), nsub = 7 leaves a ragged
    last channel group."""
    import torch
    P, L = 9, 40
    nchan = 14 if nsub != 400 else 2
    d = pkg.synth_profiles(P, L, seed=21, clouds=True)
    d["t"][4, 3] = np.nan                                   # NaN profile: NaN channel averages
    centres, bw = pkg.HATPRO_FRQ[:nchan], np.full(nchan, 0.23)
    frq = pkg.band_frequencies(nsub, centres, bw)
    w = None
    if weighted:
        w = np.random.default_rng(3).uniform(0.5, 1.5, (nchan, nsub))
        w /= w.sum(axis=1, keepdims=True)
    dev = {k: torch.from_numpy(v).cuda() for k, v in d.items()}
    args = (dev["z"], dev["p"], dev["t"], dev["rh"], dev["denliq"], dev["denice"])
    mono = pkg.Plan("R24", L, frq, pkg.HATPRO_ELEV, cloudy=True).forward(*args)
    fused = pkg.Plan("R24", L, frq, pkg.HATPRO_ELEV, cloudy=True, bandpass=(nchan, nsub, w)).forward(*args)
    torch.cuda.synchronize()
    for k in pkg.FIELDS:
        ref = cpu(pkg.bandpass(mono[k], nchan, nsub, w))
        got = cpu(fused[k])
        assert got.shape == (P, nchan, 10)
        assert np.array_equal(got, ref, equal_nan=True), k
    assert np.isnan(cpu(fused["tbtotal"])[4]).all() and np.isfinite(cpu(fused["tbtotal"])[3]).all()


@pytest.mark.parametrize("mdl", ["R98", "R17", "R20", "R24", "R22SD"])
def test_fp32_lineshape_mode(pkg, ora, mdl):
    """north_star: "any fp32 line-shape mode must stay within 0.01 K and is reported separately".
    Stage: absorption within 1e-5 relative of the fp64 oracle; end to end: TB within 0.01 K."""
    d = pkg.synth_profiles(48, 100, seed=31)
    e, _ = ora.pv.vapor(d["t"], d["rh"])
    g = pkg.absorption(mdl, d["p"], d["t"], e, pkg.HATPRO_FRQ, f32_lineshape=True)
    awet, adry, _, _ = ora.pv.absorption(mdl, d["p"], d["t"], e, pkg.HATPRO_FRQ)
    ew, ed = rel_err(cpu(g["awet"]), awet), rel_err(cpu(g["adry"]), adry)
    assert ew < 1e-5 and ed < 1e-5, (ew, ed)
    assert ew > 1e-12 or ed > 1e-12          # it really is the fp32 instantiation
    out = pkg.simulate(d["z"], d["p"], d["t"], d["rh"], pkg.HATPRO_FRQ, pkg.HATPRO_ELEV, mdl,
                       f32_lineshape=True)
    ref = ora.pv.simulate(d["z"], d["p"], d["t"], d["rh"], pkg.HATPRO_FRQ, pkg.HATPRO_ELEV, mdl)
    err = np.max(np.abs(cpu(out["tbtotal"]) - ref["tbtotal"]))
    assert err < 0.01, err


def test_fp32_lineshape_wide_spectrum(pkg, ora):
    """fp32 mode over 1-1000 GHz (far wings: large f^2 - fl^2, every cutoff class) and long lists."""
    d = pkg.synth_profiles(6, 60, seed=33)
    e, _ = ora.pv.vapor(d["t"], d["rh"])
    frq = np.concatenate([np.linspace(1.0, 1000.0, 112), [22.235, 183.31, 118.7503, 60.0]])
    g = pkg.absorption("R24", d["p"], d["t"], e, frq, f32_lineshape=True)
    awet, adry, _, _ = ora.pv.absorption("R24", d["p"], d["t"], e, frq)
    assert rel_err(cpu(g["awet"]), awet, floor=1e-12) < 2e-5
    assert rel_err(cpu(g["adry"]), adry, floor=1e-12) < 2e-5


def test_tmrcld_column(pkg, ora):
    """Upstream's `tmrcld` column (cloud mean radiating temperature) against the scalar oracle, through
    the class API (init_cloudy with cldh) and for a clear-sky object (zeros)."""
    d = pkg.synth_profiles(30, 80, seed=41, clouds=True)
    i = int(np.argmax(d["denliq"].max(axis=1)))
    wet = np.where(d["denliq"][i] > 0)[0]
    cldh = np.array([[d["z"][i][wet[0]]], [d["z"][i][wet[-1]]]])
    ang = np.array([90.0, 19.2, 4.2])
    for rt in (False, True):
        rte = pkg.TbCloudRTE(d["z"][i], d["p"][i], d["t"][i], d["rh"][i], pkg.HATPRO_FRQ, ang,
                             from_sat=False, ray_tracing=rt)
        rte.init_absmdl("R24")
        rte.init_cloudy(cldh, d["denice"][i], d["denliq"][i])
        df = rte.execute()
        o = ora.ps.TbCloudRTE(d["z"][i], d["p"][i], d["t"][i], d["rh"][i], pkg.HATPRO_FRQ, ang,
                              from_sat=False, ray_tracing=rt)
        o.init_absmdl("R24")
        o.init_cloudy(cldh, d["denice"][i], d["denliq"][i])
        ref = o.execute()
        # opaque paths below the cloud (V band at 4.2 deg: optical depth > 125 to the cloud base) give 0
        assert ((ref["tmrcld"] > 200) | (ref["tmrcld"] == 0)).all() and (ref["tmrcld"] > 200).sum() > 30
        assert np.max(np.abs(df["tmrcld"].values - ref["tmrcld"])) < TB_ABS
        assert np.max(np.abs(df["tbtotal"].values - ref["tbtotal"])) < TB_ABS
    clear = pkg.TbCloudRTE(d["z"][i], d["p"][i], d["t"][i], d["rh"][i], pkg.HATPRO_FRQ, ang, from_sat=False)
    clear.init_absmdl("R24")
    assert (clear.execute()["tmrcld"].values == 0).all()


@pytest.mark.parametrize("nf,wide", [(700, "1"), (700, "0"), (211, "1")])
def test_absorption_block_width_switch(pkg, ora, nf, wide, monkeypatch):
    """MWRT_ABS_WIDE: 28/32-wide frequency blocks (default for long lists, incl. a ragged tail) or 14/16."""
    monkeypatch.setenv("MWRT_ABS_WIDE", wide)
    d = pkg.synth_profiles(3, 40, seed=78)
    e, _ = ora.pv.vapor(d["t"], d["rh"])
    frq = np.linspace(18.0, 195.0, nf)
    g = pkg.absorption("R24", d["p"], d["t"], e, frq)
    awet, adry, _, _ = ora.pv.absorption("R24", d["p"], d["t"], e, frq)
    assert rel_err(cpu(g["awet"]), awet) < REL
    assert rel_err(cpu(g["adry"]), adry) < REL


@pytest.mark.parametrize("model", ["R24", "R98", "R22SD"])
@pytest.mark.parametrize("case", ["hatpro", "bands", "dense", "dense183", "wide_sweep", "sweep", "single",
                                  "line_centres"])
def test_far_lines_on_coarse_grid(pkg, ora, model, case, monkeypatch):
    """Far lines are summed at the block's Chebyshev nodes and interpolated (kernel 1, kNQ): the result
    must agree with the every-line-at-every-frequency form (MWRT_ABS_FAR=0) to 1e-10 and with the oracle
    to the usual gate: sub-band sets and dense spectra (28/32-wide blocks, nearly everything far, incl.
    blocks that hold line centres and the speed-dependent 183 GHz line), a wide sweep (nothing far), and
    the narrow-block kernels, which do not carry the far-line code (HATPRO, a single frequency, line
    centres)."""
    d = pkg.synth_profiles(4, 37, seed=91)
    e, _ = ora.pv.vapor(d["t"], d["rh"])
    frq = {"hatpro": pkg.HATPRO_FRQ,
           "bands": np.asarray(pkg.band_frequencies(50)).ravel()[:364],
           "dense": np.arange(49.0, 51.24, 0.01),
           "dense183": 182.0 + 0.01 * np.arange(224),
           "wide_sweep": np.linspace(1.0, 1000.0, 224),
           "sweep": np.linspace(1.0, 1000.0, 97),
           "single": np.array([31.4]),
           "line_centres": np.array([22.23508, 22.3, 60.306056, 60.4348, 118.750334, 118.8, 183.310087, 183.4,
                                     325.152888, 380.197353, 424.763, 50.0, 90.0, 150.0])}[case]
    monkeypatch.setenv("MWRT_ABS_FAR", "0")
    full = pkg.absorption(model, d["p"], d["t"], e, frq)
    monkeypatch.setenv("MWRT_ABS_FAR", "1")
    far = pkg.absorption(model, d["p"], d["t"], e, frq)
    for k in ("awet", "adry"):
        assert rel_err(cpu(far[k]), cpu(full[k])) < 1e-10, k
    awet, adry, _, _ = ora.pv.absorption(model, d["p"], d["t"], e, frq)
    assert rel_err(cpu(far["awet"]), awet) < REL
    assert rel_err(cpu(far["adry"]), adry) < REL


def test_plan_absorption_equals_stage_entry(pkg, ora):
    """mwrt_plan_absorption (kernel 1 on a plan, per-frequency-block tables cached in the plan) makes the
    same launch as the plan-less mwrt_absorption_f64 / _f32ls: bitwise equal, narrow and wide blocks,
    first call (record built) and second call (record reused)."""
    import torch
    L = 33
    d = pkg.synth_profiles(5, L, seed=3)
    e, _ = ora.pv.vapor(d["t"], d["rh"])
    p_d, t_d, e_d = (torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).cuda()
                     for x in (d["p"], d["t"], e))
    for frq in (pkg.HATPRO_FRQ, 49.0 + 0.01 * np.arange(224)):
        for f32 in (False, True):
            plan = pkg.Plan("R24", L, frq, pkg.HATPRO_ELEV[:2], f32_lineshape=f32)
            ref = pkg.absorption("R24", d["p"], d["t"], e, frq, f32_lineshape=f32)
            for _ in range(2):
                awet, adry = plan.absorption(p_d, t_d, e_d)
                torch.cuda.synchronize()
                assert torch.equal(awet, ref["awet"]) and torch.equal(adry, ref["adry"])


def test_host_entry_large_pitch_fallback(pkg, monkeypatch):
    """A 2-D copy is only used while both pitches fit cudaDeviceProp::memPitch; beyond it the host entry
    moves one field per copy.  MWRT_MAX_PITCH fakes a tiny limit; results must not change."""
    P, L = 2500, 30
    d = pkg.synth_profiles(P, L, seed=17)
    plan = pkg.Plan("R98", L, pkg.HATPRO_FRQ, pkg.HATPRO_ELEV[:4])
    a = plan.forward_host(d["z"], d["p"], d["t"], d["rh"])
    monkeypatch.setenv("MWRT_MAX_PITCH", "4096")
    b = plan.forward_host(d["z"], d["p"], d["t"], d["rh"])
    for k in pkg.FIELDS + ("status",):
        assert np.array_equal(a[k], b[k], equal_nan=True), k


def test_forward_rejects_bad_cloud_inputs(pkg):
    """Cloudy plans check denliq/denice like z/p/t/rh; the host entry checks a caller-supplied `out`."""
    import torch
    d = pkg.synth_profiles(4, 20, seed=1, clouds=True)
    dev = {k: torch.from_numpy(v).cuda() for k, v in d.items()}
    plan = pkg.Plan("R98", 20, pkg.HATPRO_FRQ, [90.0], cloudy=True)
    with pytest.raises(ValueError):
        plan.forward(dev["z"], dev["p"], dev["t"], dev["rh"])                                   # missing
    with pytest.raises(ValueError):
        plan.forward(dev["z"], dev["p"], dev["t"], dev["rh"], dev["denliq"].float(), dev["denice"])   # fp32
    with pytest.raises(ValueError):
        plan.forward(dev["z"], dev["p"], dev["t"], dev["rh"], dev["denliq"][:, :10], dev["denice"])   # shape
    with pytest.raises(ValueError):
        plan.forward(dev["z"], dev["p"], dev["t"], dev["rh"], dev["denliq"].cpu(), dev["denice"])     # host
    with pytest.raises(ValueError):
        plan.forward_host(d["z"], d["p"], d["t"], d["rh"], d["denliq"], d["denice"],
                          out={"tbtotal": np.empty((4, 14, 1), dtype=np.float32)})
    with pytest.raises(ValueError):
        plan.forward_host(d["z"], d["p"], d["t"], d["rh"], d["denliq"], d["denice"],
                          out={"tbtotal": np.empty((5, 14, 1))})


def test_host_entry_equals_device_entry_and_chunks(pkg):
    """mwrt_plan_forward_host (chunked H2D/compute/D2H pipeline) == device entry, bit for bit;
    ragged last chunk; tau_layer output."""
    import torch
    d = pkg.synth_profiles(5000, 40, seed=6)
    plan = pkg.Plan("R17", 40, pkg.HATPRO_FRQ, pkg.HATPRO_ELEV, tau_layers=True, chunk_profiles=1536)
    dev = plan.forward(*(torch.from_numpy(d[k]).cuda() for k in ("z", "p", "t", "rh")))
    torch.cuda.synchronize()
    host = plan.forward_host(d["z"], d["p"], d["t"], d["rh"])
    for k in pkg.FIELDS + ("tau_layer", "status"):
        assert np.array_equal(cpu(dev[k]), host[k], equal_nan=True), k
    assert plan.launch_count(5000) == 4 * 2   # absorption + RTE per chunk (plane-parallel)


def test_large_batch_properties(pkg):
    """Full config-2 size (10^4 x 100 x 14 x 10): size-independent properties instead of the oracle."""
    d = pkg.synth_profiles(10000, 100, seed=1234)
    out = pkg.simulate(d["z"], d["p"], d["t"], d["rh"], pkg.HATPRO_FRQ, pkg.HATPRO_ELEV, "R24")
    tb, tbatm, tmr = cpu(out["tbtotal"]), cpu(out["tbatm"]), cpu(out["tmr"])
    assert np.isfinite(tb).all()
    tmax = d["t"].max(axis=1)[:, None, None]
    tmin = d["t"].min(axis=1)[:, None, None]
    assert np.all(tb > 2.7) and np.all(tb <= tmax + 1e-6)
    assert np.all(tbatm <= tb + 1e-9)
    assert np.all(tmr >= tmin - 1e-6) and np.all(tmr <= tmax + 1e-6)
    am = 1.0 / np.sin(np.deg2rad(pkg.HATPRO_ELEV))
    tw = cpu(out["tauwet"])
    assert rel_err(tw, tw[:, :, :1] * am[None, None, :]) < 1e-12
    assert np.all(np.diff(tb[:, :7, :], axis=2) > 0)      # K band: TB rises towards the horizon
    # batch invariance: a profile's result does not depend on its neighbours
    sub = pkg.simulate(d["z"][4000:4010], d["p"][4000:4010], d["t"][4000:4010], d["rh"][4000:4010],
                       pkg.HATPRO_FRQ, pkg.HATPRO_ELEV, "R24")
    assert np.array_equal(cpu(sub["tbtotal"]), tb[4000:4010])


def test_dense_spectrum_and_icon_shapes(pkg, ora):
    """Config 4 (dense 20-60 GHz, 150 levels, per-layer tau) and config 5 (65 levels) at small P."""
    frq = pkg.dense_spectrum()[::40]       # 101 frequencies -> tail block of the 16-wide kernel
    d = pkg.synth_profiles(6, 150, seed=31)
    out = pkg.simulate(d["z"], d["p"], d["t"], d["rh"], frq, [90.0], "R24", tau_layers=True)
    ref = ora.pv.simulate(d["z"], d["p"], d["t"], d["rh"], frq, [90.0], "R24", return_layers=True)
    assert rel_err(cpu(out["tau_layer"])[:, :, 1:], ref["tau_layer"][:, :, 0, 1:]) < REL
    assert np.max(np.abs(cpu(out["tbtotal"]) - ref["tbtotal"])) < TB_ABS
    d = pkg.synth_profiles(32, 65, seed=32, top_km=22.0, exponent=2.5, lowest_layer_km=0.02)
    out = pkg.simulate(d["z"], d["p"], d["t"], d["rh"], pkg.HATPRO_FRQ, pkg.HATPRO_ELEV, "R20")
    ref = ora.pv.simulate(d["z"], d["p"], d["t"], d["rh"], pkg.HATPRO_FRQ, pkg.HATPRO_ELEV, "R20")
    assert np.max(np.abs(cpu(out["tbtotal"]) - ref["tbtotal"])) < TB_ABS


def test_dfma_peak_is_plausible(pkg):
    fl, mhz = pkg.measure_dfma_peak()
    assert 5e12 < fl < 6e13, fl      # B200 nominal 37.2 TFLOP/s fp64


def test_device_math_helpers(pkg):
    """The branch-free exp/log/reciprocal the kernels are built from, against libm, in ulps."""
    from mwr_fast_forward_operators_and_lbls_b200.forward import math_selftest
    rng = np.random.default_rng(11)

    def ulps(got, ref):
        return np.max(np.abs(got - ref) / np.spacing(np.abs(ref)))
    x = np.concatenate([rng.uniform(-700, 700, 200000), rng.uniform(-1, 1, 200000),
                        rng.uniform(-1e-3, 1e-3, 50000), [0.0, -0.0, 699.9, -699.9],
                        (np.arange(-130, 131) * np.log(2) / 64)])
    # exp_tb reduces with one constant: ~1 ulp + |x| * 1.1e-16 relative
    got = math_selftest("exp_tb", x).cpu().numpy()
    ref = np.exp(x)
    assert np.max(np.abs(got - ref) / ref / (2.3e-16 + 1.2e-16 * np.abs(x))) <= 1.0
    x = np.concatenate([np.exp(rng.uniform(-600, 600, 200000)), rng.uniform(0.5, 2.0, 200000),
                        1.0 + rng.uniform(-1e-6, 1e-6, 50000), 1.0 + rng.uniform(-0.02, 0.02, 50000),
                        [1.0, 2.0, 0.5, 1.0 + 27 / 64, np.nextafter(1.0 + 27 / 64, 0), np.nextafter(1.0, 0),
                         np.nextafter(1.0, 2), np.nextafter(2.0, 0), 1e-300, 1e300]])
    for op in ("log_tb",):
        got = math_selftest(op, x).cpu().numpy()
        ref = np.log(x)
        nz = ref != 0.0
        assert (got[~nz] == 0.0).all(), op
        assert ulps(got[nz], ref[nz]) <= 2.5, op
    x = np.concatenate([np.exp(rng.uniform(-300, 300, 200000)), -np.exp(rng.uniform(-300, 300, 1000))])
    assert ulps(math_selftest("rcp", x).cpu().numpy(), 1.0 / x) <= 1.5
    # the MUFU seed must stay good to ~2^-20 for rcp_nr1's 1e-12 claim (measured: 2^-19.97 worst case)
    seed = math_selftest("rcp_seed", x).cpu().numpy()
    assert np.max(np.abs(seed * x - 1.0)) < 2.0 ** -19.5
    nr1 = math_selftest("rcp_nr1", x).cpu().numpy()
    assert np.max(np.abs(nr1 * x - 1.0)) < 2e-12
    # branch-free atan2 (cloud liquid permittivity): all four quadrants, both |y|<|x| and |y|>|x|
    x = np.concatenate([rng.normal(size=200000) * 10.0 ** rng.uniform(-6, 6, 200000),
                        [0.0, 1.0, -1.0, 0.25, 0.75, 0.2499999, 0.7500001, 1e-300, 1e300]])
    for op, ref in (("atan2_x1", np.arctan2(x, 1.0)), ("atan2_xm1", np.arctan2(x, -1.0)),
                    ("atan2_1x", np.arctan2(1.0, x)), ("atan2_m1x", np.arctan2(-1.0, x))):
        got = math_selftest(op, x).cpu().numpy()
        nz = ref != 0.0
        assert (got[~nz] == 0.0).all(), op
        assert ulps(got[nz], ref[nz]) <= 3.0, op


def test_host_entry_field_subsets_nan_patch_and_layouts(pkg):
    """Host entry: (a) one pinned block (single 2-D copy per chunk) vs separately allocated arrays,
    (b) a subset of fields, (c) clear sky: tauliq/tauice are filled on the host -- zero, NaN for the
    flagged profile -- and must equal what the device entry writes, (d) cloudy plans move them."""
    import torch
    P, L = 3000, 30
    d = pkg.synth_profiles(P, L, seed=16, clouds=True)
    d["t"][1234, 7] = np.nan                                   # one flagged profile
    plan = pkg.Plan("R24", L, pkg.HATPRO_FRQ, pkg.HATPRO_ELEV)
    fields = pkg.FIELDS
    dev = plan.forward(*(torch.from_numpy(d[k]).cuda() for k in ("z", "p", "t", "rh")), fields=fields)
    torch.cuda.synchronize()
    blk = plan.forward_host(d["z"], d["p"], d["t"], d["rh"], fields=fields)          # pinned block
    sep = {f: np.empty((P, 14, 10)) for f in fields}                                 # pageable, separate
    sep = plan.forward_host(d["z"], d["p"], d["t"], d["rh"], fields=fields, out=sep)
    for k in fields + ("status",):
        assert np.array_equal(cpu(dev[k]), blk[k], equal_nan=True), k
        assert np.array_equal(cpu(dev[k]), sep[k], equal_nan=True), k
    assert np.isnan(blk["tauliq"][1234]).all() and blk["status"][1234] != 0
    assert (np.delete(blk["tauliq"], 1234, axis=0) == 0.0).all()
    sub = plan.forward_host(d["z"], d["p"], d["t"], d["rh"], fields=("tbtotal", "tauliq"))
    assert set(sub) == {"tbtotal", "tauliq", "status"}
    assert np.array_equal(sub["tbtotal"], blk["tbtotal"], equal_nan=True)
    assert np.array_equal(sub["tauliq"], blk["tauliq"], equal_nan=True)
    # cloudy: liquid optical depth is computed and moved
    cplan = pkg.Plan("R24", L, pkg.HATPRO_FRQ, pkg.HATPRO_ELEV, cloudy=True)
    zero = np.zeros_like(d["denliq"])
    cdev = cplan.forward(*(torch.from_numpy(d[k]).cuda() for k in ("z", "p", "t", "rh")),
                         denliq=torch.from_numpy(d["denliq"]).cuda(), denice=torch.from_numpy(zero).cuda())
    torch.cuda.synchronize()
    chost = cplan.forward_host(d["z"], d["p"], d["t"], d["rh"], denliq=d["denliq"], denice=zero)
    for k in fields:
        assert np.array_equal(cpu(cdev[k]), chost[k], equal_nan=True), k
    assert np.nanmax(chost["tauliq"]) > 0.0


def test_fused_gather_single_rank(pkg):
    """The gather fused into kernel 3 (result stores fanned out through symmetric memory, staged and
    coalesced per warp) on a one-rank group: must reproduce the plain stores bit for bit.  The real
    multi-GPU check is tools/check_fused_gather.py under torchrun (profiles/r01j_scaling.txt)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, MWRT_CHECK_P="777", MWRT_CHECK_NOTIME="1")
    import socket
    for mode in ("peers", "auto"):
        with socket.socket() as sk:          # a free port per launch
            sk.bind(("127.0.0.1", 0))
            port = sk.getsockname()[1]
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "1",
                            "--master-addr", "127.0.0.1", "--master-port", str(port),
                            os.path.join(root, "tools", "check_fused_gather.py"), mode],
                           capture_output=True, text=True, timeout=300, env=env, cwd=root)
        if r.returncode == 77:
            pytest.skip("symmetric memory is not available on this box: " + r.stdout[-300:])
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        assert "identical on every rank: True" in r.stdout


@pytest.mark.parametrize("nf", [112, 128, 130, 140])
def test_absorption_wide_frequency_blocks(pkg, ora, nf):
    """Long frequency lists take the 28- / 32-wide kernels (incl. a ragged tail block at F=130)."""
    d = pkg.synth_profiles(5, 50, seed=77)
    e, _ = ora.pv.vapor(d["t"], d["rh"])
    frq = np.linspace(18.0, 195.0, nf)
    g = pkg.absorption("R22SD", d["p"], d["t"], e, frq)
    awet, adry, _, _ = ora.pv.absorption("R22SD", d["p"], d["t"], e, frq)
    assert rel_err(cpu(g["awet"]), awet) < REL
    assert rel_err(cpu(g["adry"]), adry) < REL


def test_no_out_of_bounds_writes_guard_bands(pkg):
    """compute-sanitizer is closed on this GPU pool (profiles/r02_sanitizer_closed.log), so out-of-bounds
    WRITES are hunted with guard bands instead: every output of every entry point is a window of a larger
    buffer pre-filled with a sentinel bit pattern; after the kernels ran on awkward sizes (ragged tiles,
    ragged warps, ragged channel groups) the bands before and after each window must be untouched and
    the window itself fully overwritten."""
    import torch
    from mwr_fast_forward_operators_and_lbls_b200 import _lib
    from mwr_fast_forward_operators_and_lbls_b200.forward import _ptr, _stream
    lib = _lib.load()
    SENT = -1.2345678912345e+200
    G = 257                                                    # guard elements either side

    def guarded(shape, dtype=torch.float64):
        n = int(np.prod(shape))
        buf = torch.full((n + 2 * G,), SENT if dtype == torch.float64 else -77, dtype=dtype, device="cuda")
        return buf, buf[G:G + n].view(*shape)

    def check(buf, n, name, must_fill=True):
        sent = SENT if buf.dtype == torch.float64 else -77
        assert bool((buf[:G] == sent).all()) and bool((buf[G + n:] == sent).all()), "guard band hit: " + name
        if must_fill:
            assert not bool((buf[G:G + n] == sent).any()), "window not fully written: " + name

    P, L = 37, 33
    d = pkg.synth_profiles(P, L, seed=5, clouds=True)
    dev = {k: torch.from_numpy(v).cuda() for k, v in d.items()}
    cases = [("clear", dict(), pkg.HATPRO_FRQ, pkg.HATPRO_ELEV, None),
             ("cloudy+rt", dict(cloudy=True, ray_tracing=True), pkg.HATPRO_FRQ[:5], pkg.HATPRO_ELEV[:3], None),
             ("bandpass 14x5", dict(cloudy=True), pkg.band_frequencies(5), pkg.HATPRO_ELEV, (14, 5, None)),
             ("bandpass 3x7 ragged groups", dict(), pkg.band_frequencies(7, pkg.HATPRO_FRQ[:3], np.full(3, 0.2)),
              pkg.HATPRO_ELEV[:7], (3, 7, None)),
             ("tau_layers 130 freqs", dict(tau_layers=True), np.linspace(20, 60, 130), np.array([90.0, 30.0]), None),
             ("fp32 line shape", dict(f32_lineshape=True), pkg.HATPRO_FRQ, pkg.HATPRO_ELEV[:4], None)]
    for name, kw, frq, ang, bp in cases:
        plan = pkg.Plan("R24", L, frq, ang, bandpass=bp, **kw)
        bufs, out = {}, {}
        for f in pkg.FIELDS:
            bufs[f], out[f] = guarded((P, plan.Fo, plan.A))
        bufs["status"], out["status"] = guarded((P,), torch.int32)
        if kw.get("tau_layers"):
            bufs["tau_layer"], out["tau_layer"] = guarded((P, plan.F, L))
        plan.forward(dev["z"], dev["p"], dev["t"], dev["rh"], dev["denliq"] if plan.cloudy else None,
                     dev["denice"] if plan.cloudy else None, out=out)
        torch.cuda.synchronize()
        for k, b in bufs.items():
            check(b, out[k].numel(), "%s/%s" % (name, k))
    # stage entry points
    e, _ = pkg.vapor(dev["t"], dev["rh"])
    frq = torch.from_numpy(np.ascontiguousarray(pkg.HATPRO_FRQ[:13])).cuda()
    m = pkg.get_model("R22SD")
    F = 13
    b_w, awet = guarded((P, F, L))
    b_d, adry = guarded((P, F, L))
    b_l, aliq = guarded((P, F, L))
    b_i, aice = guarded((P, F, L))
    b_s, st = guarded((P,), torch.int32)
    st.zero_()
    _lib.check(lib.mwrt_absorption_f64(m.handle, P, L, F, _ptr(dev["p"]), _ptr(dev["t"]), _ptr(e), _ptr(dev["denliq"]),
                                       _ptr(dev["denice"]), _ptr(frq), _ptr(awet), _ptr(adry), _ptr(aliq),
                                       _ptr(aice), _ptr(st), _stream()), "absorption")
    ang = torch.from_numpy(np.ascontiguousarray(pkg.HATPRO_ELEV[:7])).cuda()
    b_ds, ds = guarded((P, 7, L))
    _lib.check(lib.mwrt_slant_path_f64(P, L, 7, _ptr(dev["z"]), _ptr(dev["p"]), _ptr(dev["t"]), _ptr(e), _ptr(ang),
                                       1, _ptr(ds), _ptr(st), _stream()), "slant_path")
    torch.cuda.synchronize()
    for b, n, nm in ((b_w, awet.numel(), "awet"), (b_d, adry.numel(), "adry"), (b_l, aliq.numel(), "aliq"),
                     (b_i, aice.numel(), "aice"), (b_ds, ds.numel(), "ds"), (b_s, P, "status")):
        check(b, n, "stage/" + nm)


@pytest.mark.parametrize("P,L,F,A", [(1, 2, 1, 1), (3, 5, 2, 3), (2, 300, 3, 11), (1, 50, 17, 20),
                                     (129, 33, 15, 7), (5, 100, 29, 2)])
@pytest.mark.parametrize("mode", ["plane", "raytraced", "cloudy", "upwelling"])
def test_edge_shapes_against_oracle(pkg, ora, P, L, F, A, mode):
    """Smallest and most awkward shapes of every variant: 2 levels (one layer), one frequency / angle,
    more angles than a register tile holds (11, 20 -> two angle groups), frequency counts that are no
    multiple of any block width, level counts beyond a warp round of the slant path (300), tile- and
    warp-ragged batches (129 x 33)."""
    Lg = max(L, 5)                                         # (the generator needs a few levels to smooth over)
    d = pkg.synth_profiles(P, Lg, seed=100 + P + L, clouds=(mode == "cloudy"))
    pick = np.unique(np.linspace(0, Lg - 1, L).round().astype(int))
    d = {k: np.ascontiguousarray(v[:, pick]) for k, v in d.items()}
    frq = np.linspace(20.0, 61.0, F) if F > 1 else np.array([31.4])
    ang = np.linspace(90.0, 4.2, A) if A > 1 else np.array([19.2])
    if mode == "raytraced":
        ang = np.clip(ang, 4.2, 88.5)                      # (upstream takes 89-91 deg as the vertical)
    kw = dict(ray_tracing=(mode == "raytraced"), from_sat=(mode == "upwelling"))
    if mode == "cloudy":
        kw.update(denliq=d["denliq"], denice=np.where(d["t"] < 250.0, 0.01, 0.0))
    out = pkg.simulate(d["z"], d["p"], d["t"], d["rh"], frq, ang, "R24", **kw)
    ref = ora.pv.simulate(d["z"], d["p"], d["t"], d["rh"], frq, ang, "R24", **kw)
    for k in ("tbtotal", "tbatm", "tmr"):
        assert np.max(np.abs(cpu(out[k]) - ref[k])) < TB_ABS, k
    for k in ("tauwet", "taudry", "tauliq", "tauice"):
        assert rel_err(cpu(out[k]), ref[k], floor=1e-30) < REL, k
    assert (cpu(out["status"]) == 0).all()


def test_physics_invariants_on_the_gpu(pkg):
    """SURVEY.md Appendix D on the CUDA path itself: dry air -> tau_wet = 0 exactly; no cloud water ->
    tau_liq = tau_ice = 0 exactly on a cloudy plan; an almost empty atmosphere -> TB -> the cosmic
    background (Planck-consistent, 2.736 K); results do not depend (beyond rounding) on the other angles
    or frequencies a profile is evaluated with."""
    import torch
    d = pkg.synth_profiles(6, 60, seed=9)
    zero = np.zeros_like(d["rh"])
    out = pkg.simulate(d["z"], d["p"], d["t"], zero, pkg.HATPRO_FRQ, pkg.HATPRO_ELEV, "R24")
    assert (cpu(out["tauwet"]) == 0.0).all() and np.isfinite(cpu(out["tbtotal"])).all()
    out = pkg.simulate(d["z"], d["p"], d["t"], d["rh"], pkg.HATPRO_FRQ, pkg.HATPRO_ELEV, "R24",
                       denliq=zero, denice=zero)
    clr = pkg.simulate(d["z"], d["p"], d["t"], d["rh"], pkg.HATPRO_FRQ, pkg.HATPRO_ELEV, "R24")
    assert (cpu(out["tauliq"]) == 0.0).all() and (cpu(out["tauice"]) == 0.0).all()
    assert np.array_equal(cpu(out["tbtotal"]), cpu(clr["tbtotal"]))
    thin = pkg.simulate(d["z"], d["p"] * 1e-7, d["t"], zero, [31.4], [90.0], "R24")
    assert np.max(np.abs(cpu(thin["tbtotal"]) - 2.736)) < 1e-6
    one = pkg.simulate(d["z"], d["p"], d["t"], d["rh"], pkg.HATPRO_FRQ[3:4], pkg.HATPRO_ELEV[5:6], "R24")
    # (a different frequency-block width and angle tile compile to different instruction orders:
    # agreement to rounding, not bit for bit)
    assert np.max(np.abs(cpu(one["tbtotal"])[:, 0, 0] - cpu(clr["tbtotal"])[:, 3, 5])) < 1e-9
    del torch
