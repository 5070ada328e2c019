"""CPU tests of the oracle itself (no GPU): the PyRTlib-shaped scalar restatement against the
independently written vectorised one, against the committed golden vectors, and against the physics
invariants of SURVEY.md Appendix D.  The reference has no tests or goldens for this path
 -- parity with real PyRTlib is unpinned.
"""
import json
import os

import numpy as np
import pytest

from oracle import atmospheres, model_tables, pyrtlib_scalar as ps, pyrtlib_vec as pv

FRQ = np.array([22.24, 23.04, 23.84, 25.44, 26.24, 27.84, 31.4, 51.26, 52.28, 53.86, 54.94, 56.66,
                57.3, 58.0])
ANG = np.array([90., 30, 19.2, 14.4, 11.4, 8.4, 6.6, 5.4, 4.8, 4.2])
GOLD = os.path.join(os.path.dirname(__file__), "golden", "config1_us_standard.json")


def _rel(a, b):
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))


def test_models_the_reference_uses_exist():
    have = model_tables.implemented_models()
    # declared at PyRTlib_processing.py:90, run at :122,130,138,146; R22SD from BASELINE.json config 1
    for m in ["R17", "R03", "R16", "R19", "R98", "R19SD", "R20", "R20SD", "R24", "R22SD"]:
        assert m in have
    with pytest.raises(ValueError):
        model_tables.load("R99")


@pytest.mark.parametrize("mdl", ["R98", "R17", "R19SD", "R22SD"])
@pytest.mark.parametrize("rt", [False, True])
def test_scalar_and_vector_restatements_agree(mdl, rt):
    z, p, t, rh = atmospheres.us_standard()
    ang = ANG[[0, 3, 9]]
    r = ps.TbCloudRTE(z, p, t, rh, FRQ, ang, ray_tracing=rt)
    r.init_absmdl(mdl)
    r.satellite = False
    d = r.execute()
    o = pv.simulate(z[None], p[None], t[None], rh[None], FRQ, ang, mdl, ray_tracing=rt)
    for k in ("tbtotal", "tbatm", "tmr", "tauwet", "taudry"):
        a = d[k].reshape(len(ang), len(FRQ)).T
        assert _rel(o[k][0], a) < 1e-12, k


def test_scalar_and_vector_agree_cloudy_and_118ghz():
    z, p, t, rh = atmospheres.us_standard()
    nl = 30
    z, p, t, rh = z[:nl], p[:nl], t[:nl], rh[:nl]
    denliq = np.zeros(nl)
    denliq[2:4] = [0.2, 0.3]
    denice = np.zeros(nl)
    denice[8:10] = 0.05
    frq = np.array([22.24, 31.4, 90.0, 118.0, 118.75, 119.5, 183.0])
    for mdl in ("R98", "R20SD"):
        r = ps.TbCloudRTE(z, p, t, rh, frq, [90.0, 19.2])
        r.init_absmdl(mdl)
        r.satellite = False
        r.init_cloudy(np.array([[2.0], [3.0]]), denice, denliq)
        d = r.execute()
        o = pv.simulate(z[None], p[None], t[None], rh[None], frq, [90.0, 19.2], mdl, denliq=denliq[None],
                        denice=denice[None])
        for k in ("tbtotal", "tauliq", "tauice", "taudry", "tauwet"):
            a = d[k].reshape(2, len(frq)).T
            assert _rel(o[k][0], a) < 1e-12, (mdl, k)
        assert np.all(o["tauliq"] > 0) and np.all(o["tauice"] > 0)


def test_golden_vectors_regression():
    g = json.load(open(GOLD))
    z, p, t, rh = (np.array(g[k]) for k in ("z", "p", "t", "rh"))
    frq, ang = np.array(g["frq"]), np.array(g["angles"])
    for case in g["cases"]:
        o = pv.simulate(z[None], p[None], t[None], rh[None], frq, ang, case["model"],
                        ray_tracing=case["ray_tracing"])
        for k in ("tbtotal", "tbatm", "tmr", "tauwet", "taudry"):
            a = np.array(case[k]).reshape(len(ang), len(frq)).T
            assert _rel(o[k][0], a) < 1e-11, (case["model"], k)


def test_upwelling_matches_between_restatements():
    z, p, t, rh = atmospheres.us_standard()
    r = ps.TbCloudRTE(z, p, t, rh, FRQ, [90.0, 30.0])
    r.init_absmdl("R19")
    assert r.satellite is True  # upstream default from_sat=True; the reference overrides it (:125)
    d = r.execute()
    o = pv.simulate(z[None], p[None], t[None], rh[None], FRQ, [90.0, 30.0], "R19", from_sat=True)
    a = d["tbtotal"].reshape(2, len(FRQ)).T
    assert _rel(o["tbtotal"][0], a) < 1e-12


def test_physics_invariants_us_standard():
    """SURVEY.md Appendix D."""
    z, p, t, rh = atmospheres.us_standard()
    for mdl in ("R98", "R17", "R20", "R24"):
        o = pv.simulate(z[None], p[None], t[None], rh[None], FRQ, ANG, mdl, return_layers=True)
        tb = o["tbtotal"][0]
        assert np.all(tb > 0) and np.all(tb <= t.max() + 1e-6)
        assert np.all(o["tbatm"][0] <= tb + 1e-9)
        assert np.all(o["tmr"][0] >= t.min() - 1e-6) and np.all(o["tmr"][0] <= t.max() + 1e-6)
        # tau(angle) = tau(zenith)/sin(elev) in plane-parallel mode
        am = 1.0 / np.sin(np.deg2rad(ANG))
        assert _rel(o["tauwet"][0], o["tauwet"][0][:, :1] * am[None, :]) < 1e-12
        # opaque V-band ~ near-surface temperature; transparent K-band TB grows as elevation drops
        assert np.all(np.abs(tb[11:, :] - t[0]) < 2.5)
        assert np.all(np.diff(tb[:7, :], axis=1) > 0)
        # zenith magnitudes (order-of-magnitude sanity windows)
        zen = tb[:, 0]
        assert np.all((zen[:7] > 12) & (zen[:7] < 40))
        assert 95 < zen[7] < 120 and 135 < zen[8] < 160 and 235 < zen[9] < 258
        assert np.all(zen[10:] > 275)
        # absorption identical across the angle axis is structural here (computed once)
        assert o["awet"].shape == (1, len(FRQ), len(z))


def test_dry_atmosphere_and_zero_liquid():
    z, p, t, rh = atmospheres.us_standard()
    o = pv.simulate(z[None], p[None], t[None], 0 * rh[None], FRQ, [90.0], "R20")
    assert np.all(o["tauwet"] == 0.0)
    o2 = pv.simulate(z[None], p[None], t[None], rh[None], FRQ, [90.0], "R20", denliq=0 * z[None])
    assert np.all(o2["tauliq"] == 0.0)


def test_ray_traced_path_shorter_than_plane_parallel():
    z, p, t, rh = atmospheres.us_standard()
    e, _ = pv.vapor(t, rh)
    ds_pp = pv.slant_path(z[None], p[None], t[None], e[None], ANG, False)[0]
    ds_rt = pv.slant_path(z[None], p[None], t[None], e[None], ANG, True)[0]
    assert np.all(ds_rt[:, 1:] <= ds_pp[:, 1:] * (1 + 1e-9))
    assert np.allclose(ds_rt[0], ds_pp[0])           # 90 deg: plain height differences
    # air mass at 4.2 deg: close to 1/sin in the lowest 5 km, much less over the whole 120 km column
    assert 11.0 < ds_rt[9, 1:6].sum() / ds_pp[0, 1:6].sum() < 13.65
    assert 5.0 < ds_rt[9].sum() / ds_pp[0].sum() < 13.65


def test_nan_profile_gives_nan_and_batch_continues():
    """Reference gate: PyRTlib_processing.py:71-79,117-119,153-154."""
    z, p, t, rh = atmospheres.us_standard()
    Z, P_, T, RH = (np.stack([v, v, v]) for v in (z, p, t, rh))
    T[1, 7] = np.nan
    o = pv.simulate(Z, P_, T, RH, FRQ, ANG[:2], "R98")
    assert np.isnan(o["tbtotal"][1]).all() and np.isnan(o["tauwet"][1]).all()
    assert np.isfinite(o["tbtotal"][0]).all() and np.array_equal(o["tbtotal"][0], o["tbtotal"][2])


def test_negative_absorption_raises_like_upstream():
    with pytest.raises(ValueError):
        ps.exponential_integration(True, np.array([1.0, -1.0, 2.0]), np.array([0.0, 1.0, 1.0]), 0, 3, 1)


def test_specific_attenuation_matches_textbook_values():
    """External sanity anchor for the from-memory line tables (NOT parity): at sea level (1013.25 hPa,
    15 C, 7.5 g m^-3 of vapour) the restatement gives the specific attenuations every microwave
    propagation text quotes for that state (ITU-R P.676-class curves): ~0.18 dB/km at the 22.235 GHz
    water line, ~29 dB/km at 183.31 GHz, ~15 dB/km in the 60 GHz oxygen complex, ~1.3 dB/km at the
    118.75 GHz oxygen line, ~0.4 dB/km of vapour at 94 GHz.  12 % windows: the point is to catch a
    mis-remembered strength, width or unit, not to validate digits."""
    from oracle import pyrtlib_vec as pv
    p, t = np.array([[1013.25]]), np.array([[288.15]])
    e = np.array([[7.5e-3 * 461.5 * 288.15 / 100.0]])          # 7.5 g m^-3 -> mb
    frq = np.array([22.235, 183.31, 60.0, 118.75, 94.0])
    db = 10.0 / np.log(10.0)
    expect_wet = {0: 0.18, 1: 29.0, 4: 0.38}
    expect_dry = {2: 14.8, 3: 1.33}
    for mdl in ("R98", "R17", "R20", "R24", "R22SD"):
        awet, adry, _, _ = pv.absorption(mdl, p, t, e, frq)
        for j, v in expect_wet.items():
            assert abs(awet[0, j, 0] * db - v) < 0.12 * v, (mdl, frq[j], awet[0, j, 0] * db)
        for j, v in expect_dry.items():
            assert abs(adry[0, j, 0] * db - v) < 0.12 * v, (mdl, frq[j], adry[0, j, 0] * db)
