"""Live cross-checks of the oracle ports against the reference's OWN functions, imported from
/root/reference in this container (with the stubs of tests/golden/make_*_golden.py for the modules the
files import but this image lacks).  Skipped wherever /root/reference does not exist (the GPU box);
the committed goldens (tests/golden/*.npz) cover the same functions there.  Fresh seeds, so these
exercise cases the goldens do not hold."""
import importlib.util
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/python_src"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(HERE, "golden", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="module")
def ref_modules():
    """The reference's preprocessing4all.py and derive_cloud_water.py, imported once.  The stubs go
    into sys.modules only for the duration of the import and are removed again."""
    saved = dict(sys.modules)
    saved_path = list(sys.path)
    try:
        gp, gc = _load("make_preproc_golden"), _load("make_cloud_golden")
        gp._install_stubs()
        del sys.modules["derive_cloud_water"]       # the real one imports fine with the other stubs
        sys.path.insert(0, os.path.join(REF, "preproc"))
        import derive_cloud_water as cld           # noqa: E402
        import preprocessing4all as pre            # noqa: E402
        yield gp, gc, pre, cld
    finally:
        for k in list(sys.modules):
            if k not in saved:
                del sys.modules[k]
        sys.modules.update(saved)
        sys.path[:] = saved_path


def test_preproc_oracle_vs_live_reference(ref_modules):
    from oracle import preproc as op
    from mwr_fast_forward_operators_and_lbls_b200 import preproc as prod
    gp, _, pre, _ = ref_modules
    gp._CLIM.update(prod.midlatitude_summer())
    clim = dict(gp._CLIM)
    rng = np.random.default_rng(987)
    n_checked = 0
    for c in range(12):
        kw = dict(n=int(rng.integers(1500, 5000)), top_km=float(rng.uniform(14, 33)),
                  frac_rh=bool(c % 4 == 1), nan_frac=float(rng.uniform(0, 0.05)), hold=int(rng.integers(1, 6)))
        h, t, p, rh = gp.synth_sonde(rng, **kw)
        fname = "/x/other/live%d.nc" % c
        gp._CURRENT[fname] = gp._FakeDS(Height=gp._Var(h), Temperature=gp._Var(t), Pressure=gp._Var(p),
                                        Humidity=gp._Var(rh), Latitude=gp._Var([50.9]), Longitude=gp._Var([6.4]))
        r = pre.read_radiosonde_nc_arms(file=fname, crop=0)
        inds = op.thinning_indices(h, p)
        mine = op.assemble_levels(h, t, p, rh, inds, clim)
        if r[0] == 0:
            assert not mine["ok"]
            continue
        p2, t2, q2, m2, z2, r2 = pre.add_clim2profiles(np.array(r[1]), np.array(r[2]), np.array(r[3]),
                                                       np.array(r[6]), np.array(r[7]), np.array(r[8]))
        n = pre.n_levels
        if len(p2) < n:
            continue
        want = np.stack([p2[-n:], t2[-n:], q2[-n:], 1000 * m2[-n:], z2[-n:], r2[-n:]])
        got = np.stack([mine[k] for k in ("p", "t", "ppmv", "h2o_gkg", "z", "rh")])
        assert np.array_equal(got, want, equal_nan=True), c
        n_checked += 1
    assert n_checked >= 8


def test_cloud_oracle_vs_live_reference(ref_modules):
    from oracle import cloud
    _, gc, _, cld = ref_modules
    rng = np.random.default_rng(4242)
    hits = 0
    for c in range(60):
        p, t, z, rh = gc.synth(rng, layers=int(rng.integers(0, 6)), warm=bool(c % 2))
        with np.errstate(all="ignore"):
            want = cld.derive_cloud_features(p, t, None, None, z, rh)
        got = cloud.derive_cloud_features(p, t, z, rh)
        for k in (0, 1, 3, 4):
            assert np.array_equal(got[k], want[k], equal_nan=True), (c, k)
        assert got[2] == want[2] and got[5] == want[5]
        hits += bool((want[0] > 0).any() or (want[3] > 0).any())
    assert hits >= 20


def test_stats_oracle_vs_live_reference():
    import ast
    from oracle import postproc
    src = os.path.join(REF, "plot_scripts", "multi_campaign_plots_and_ana.py")
    fn = next(n for n in ast.parse(open(src).read()).body
              if isinstance(n, ast.FunctionDef) and n.name == "stats_by_channel")
    ns = {"np": np, "n_chans": 14}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), src, "exec"), ns)
    rng = np.random.default_rng(31)
    for T in (5, 60, 521):
        v = rng.normal(120, 50, (T, 14))
        r = v + rng.normal(-0.2, 2.0, (T, 14))
        v[rng.uniform(size=v.shape) < 0.1] = np.nan
        r[rng.uniform(size=r.shape) < 0.1] = np.nan
        with np.errstate(all="ignore"):
            want = ns["stats_by_channel"](v, r, 14)
            got = postproc.stats_by_channel(v, r, 14)
        for a, b in zip(got, want):
            assert np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)
