"""CPU checks of the algebra the CUDA kernels rest on (csrc/mwrt_absorption.cu, mwrt_rte_kernel.cuh):
the identities are replayed in NumPy against the textbook / oracle forms, so the reformulations the
kernels use for speed are verified independently of any GPU run."""
import cmath

import numpy as np

from oracle import pyrtlib_scalar as ps

rng = np.random.default_rng(12)


def test_h2o_two_wings_are_one_fraction():
    """sp*w*(1/D0 + 1/D1) = (n1*vv + n0)/(vv^2 + c4), vv = f^2 - (fc^2 - w^2)."""
    fc, w, sp = rng.uniform(20, 1000, 5000), 10 ** rng.uniform(-5, 0.7, 5000), rng.uniform(0.1, 5, 5000)
    f = rng.uniform(1, 200, 5000)
    ref = sp * (w / ((f - fc) ** 2 + w * w) + w / ((f + fc) ** 2 + w * w))
    fc2, wsq = fc * fc, w * w
    c3, c4, n1 = fc2 - wsq, 4 * fc2 * wsq, 2 * sp * w
    n0 = n1 * (2 * fc2)
    vv = f * f - c3
    got = (n1 * vv + n0) / (vv * vv + c4)
    far = np.abs(f - fc) > 1e-3 * fc           # the kernel corrects narrow lines close to a frequency
    assert np.max(np.abs(got[far] - ref[far]) / ref[far]) < 1e-11


def test_o2_line_with_mixing_is_one_fraction():
    """sp*[(df*g + d1*y)/(d1^2+df^2) + (df*g - d2*y)/(d2^2+df^2)] with d1 = f - fc, d2 = f + fc."""
    n = 5000
    fc, df, sp = rng.uniform(48, 70, n), 10 ** rng.uniform(-4, 0.4, n), rng.uniform(0.1, 5, n)
    g, y = 1 + rng.uniform(-0.05, 0.05, n), rng.uniform(-0.6, 0.6, n)
    f = rng.uniform(20, 60, n)
    d1, d2 = f - fc, f + fc
    ref = sp * ((df * g + d1 * y) / (d1 * d1 + df * df) + (df * g - d2 * y) / (d2 * d2 + df * df))
    As, P, fc2, df2 = sp * df * g, sp * y * fc, fc * fc, df * df
    c3, c4 = fc2 - df2, 4 * fc2 * df2
    n1, n0 = 2 * (As + P), 4 * (As * fc2 - P * df2)
    vv = f * f - c3
    got = (n1 * vv + n0) / (vv * vv + c4)
    far = np.abs(d1) > 1e-3 * fc
    scale = sp * (df * np.abs(g) + np.abs(d1 * y)) / (d1 * d1 + df * df)      # size of the terms
    assert np.max(np.abs(got[far] - ref[far]) / scale[far]) < 1e-10


def test_two_lines_share_one_reciprocal():
    na, nb, Da, Db = (rng.uniform(-3, 3, 1000), rng.uniform(-3, 3, 1000), 10 ** rng.uniform(-8, 12, 1000),
                      10 ** rng.uniform(-8, 12, 1000))
    ref = na / Da + nb / Db
    got = (nb * Da + na * Db) / (Da * Db)
    assert np.max(np.abs(got - ref) / (np.abs(na / Da) + np.abs(nb / Db))) < 1e-15


def _sd_fast(w0, w2, d2, df):
    """sd_shape_fast of csrc/mwrt_absorption.cu, replayed: hoisted 1/(w2 - i d2), real-coefficient
    quadratic-factor recurrences for Hui's P6 (times sqrt(pi)) and monic Q7."""
    q = 1.0 / (w2 * w2 + d2 * d2)
    icr, ici = w2 * q, d2 * q
    A, B0 = w0 - 1.5 * w2, 1.5 * d2
    a_re, a_im = A * icr - B0 * ici, A * ici + B0 * icr
    xr, xi = a_re - df * ici, a_im + df * icr
    m = np.sqrt(xr * xr + xi * xi)
    t = np.sqrt(0.5 * m + 0.5 * abs(xr))
    o = xi * (0.5 / t)
    zr, zi = (t, o) if xr >= 0 else (abs(o), np.copysign(t, xi))
    r = 2 * zr
    a = [ps.SQRT_PI * c for c in ps._HUI_A]
    b = list(ps._HUI_B)
    p2, p1 = a[6], r * a[6] + a[5]
    for k in (4, 3, 2, 1, 0):
        p0 = a[k] + r * p1 - m * p2
        if k:
            p2, p1 = p1, p0
    pre, pim = p0 - p1 * zr, p1 * zi
    q2, q1 = 1.0, r + b[6]
    q0 = r * q1 + b[5] - m
    q2, q1 = q1, q0
    for k in (4, 3, 2, 1, 0):
        q0 = b[k] + r * q1 - m * q2
        if k:
            q2, q1 = q1, q0
    qre, qim = q0 - q1 * zr, q1 * zi
    rd = 1.0 / (qre * qre + qim * qim)
    wre, wim = (pre * qre + pim * qim) * rd, (pim * qre - pre * qim) * rd
    pxr, pxi = zr * wre - zi * wim, zr * wim + zi * wre
    return pxi * (2 * ici) + (1.0 - pxr) * (2 * icr)


def test_speed_dependent_shape_reformulation():
    """The restructured speed-dependent shape equals the oracle's complex-arithmetic form."""
    worst = 0.0
    for _ in range(3000):
        w0 = 10 ** rng.uniform(-3, 0.5)
        w2 = w0 * rng.uniform(0.05, 0.3)
        d2 = w0 * rng.uniform(-0.05, 0.05)
        df = rng.uniform(-10, 10) * w0
        cden = complex(w2, -d2)
        xrt = cmath.sqrt(complex(w0 - 1.5 * w2, df + 1.5 * d2) / cden)
        pxw = ps.SQRT_PI * xrt * ps.dcerror(-xrt.imag, xrt.real)
        ref = (2.0 * (1.0 - pxw) / cden).real
        got = _sd_fast(w0, w2, d2, df)
        worst = max(worst, abs(got - ref) / abs(ref))
    assert worst < 1e-9      # 1 - sqrt(pi) z w(z) cancels up to ~1e3: well inside the 1e-9 budget


def test_layer_weight_as_transmission_difference():
    """B*(1 - e^-tau)*T_above accumulated as B*(T_before - T_after), T carried as a running product."""
    tau = 10 ** rng.uniform(-6, 1, (200, 99))
    B = rng.uniform(0.5, 2.0, (200, 99))
    ex = np.exp(-tau)
    ref = np.sum(B * (1 - ex) * np.exp(-np.concatenate([np.zeros((200, 1)), np.cumsum(tau, 1)[:, :-1]], 1)), 1)
    trans, acc = np.ones(200), np.zeros(200)
    for l in range(99):
        tn = trans * ex[:, l]
        acc += B[:, l] * (trans - tn)
        trans = tn
    assert np.max(np.abs(acc - ref) / ref) < 1e-12


def test_single_newton_reciprocal_budget():
    """rcp_nr1: a seed good to 2^-19.97 (measured on B200) gives 1/x to 1e-12 after one quadratic step."""
    x = 10 ** rng.uniform(-12, 24, 100000)
    seed = (1.0 / x) * (1.0 + rng.uniform(-1, 1, x.size) * 2.0 ** -19.97)
    r = seed + seed * (1.0 - x * seed)
    assert np.max(np.abs(r * x - 1.0)) < 1e-12


# --------------------------------------------------------------------------------------------------
# round 2: trig-free slant path, real-arithmetic cloud liquid, fp32 pair form
# --------------------------------------------------------------------------------------------------
def _asin_small(x):
    x2 = x * x
    return x * (1 + x2 * (1 / 6 + x2 * (3 / 40 + x2 * (15 / 336 + x2 * (105 / 3456 + x2 * (945 / 42240))))))


def _sin_small(x):
    x2 = x * x
    return x * (1 - x2 * (1 / 6 - x2 * (1 / 120 - x2 * (1 / 5040 - x2 / 362880))))


def _slant_kernel_algebra(z, n, angle, z0):
    """k_slant_path's per-level formulas (csrc/mwrt_rte.cu) in NumPy: no asin/tan/cos of theta, the
    bending angle as a prefix sum (np.cumsum stands in for the warp scan)."""
    L, re = len(z), ps.EARTH_RADIUS
    th0 = angle * np.pi / 180
    rs = re + z[0] + z0
    cos0, sin0, sina = np.cos(th0), np.sin(th0), np.sin(th0 * 0.5)
    a0 = 2 * sina * sina
    c40, s40 = np.cos(th0 * 0.25), np.sin(th0 * 0.25)
    i = np.arange(1, L)
    r, rl, ni, nl = re + z[i] + z0, re + z[i - 1] + z0, n[i], n[i - 1]
    same = (ni == nl) | (ni == 1.0) | (nl == 1.0)
    with np.errstate(all="ignore"):
        refbar = np.where(same, (ni + nl) * 0.5, 1.0 + (nl - ni) / np.log((nl - 1.0) / (ni - 1.0)))
    argdth = z[i] / rs - (n[0] - ni) * cos0 / ni
    argth = 0.5 * (a0 + argdth) / r
    ok = argth > 0
    sint = np.sqrt(np.where(ok, r * argth, 0.25))
    small = sint <= sin0
    cosal = np.sqrt(np.maximum(1 - sint * sint, 0.0))
    ch = np.sqrt(0.5 * (1 + cosal))
    sh_ = sint / (2 * ch)
    sind4 = (0.5 * argdth - z[i] * argth) / (2 * (sint + sina) * (ch * c40 - sh_ * s40))
    dth_s = 4 * np.where(np.abs(sind4) < 0.05, _asin_small(sind4), np.arcsin(np.clip(sind4, -1, 1)))
    c4 = np.sqrt(np.maximum(1 - sind4 * sind4, 0.0))
    s2, c2 = 2 * sind4 * c4, 1 - 2 * sind4 * sind4
    s1, c1 = 2 * s2 * c2, 1 - 2 * s2 * s2
    cot_s = (cos0 * c1 - sin0 * s1) / (sin0 * c1 + cos0 * s1)
    dth_l = 2 * np.arcsin(np.clip(sint, -1, 1)) - th0
    with np.errstate(all="ignore"):
        cot_l = (1 - 2 * sint * sint) / (2 * sint * cosal)
    dtheta, cot = np.where(small, dth_s, dth_l), np.where(small, cot_s, cot_l)
    cotl = np.concatenate([[cos0 / sin0], cot[:-1]])
    dtau = (cot + cotl) * 0.5 * (nl - ni) / refbar
    tau = np.cumsum(dtau)
    phi = dtheta + tau
    x = (phi - np.concatenate([[0.0], phi[:-1]])) * 0.5
    shx = np.where(np.abs(x) < 0.05, _sin_small(x), np.sin(x))
    dz = z[i] - z[i - 1]
    ds = np.sqrt(dz * dz + 4 * r * rl * shx * shx)
    y = np.abs(dtau) * 0.5
    with np.errstate(all="ignore"):
        corr = np.where(y < 0.05, y / _sin_small(np.where(y == 0, 1.0, y)), y / np.sin(y))
    ds = np.where(dtau != 0, ds * corr, ds)
    bad = np.where(~ok)[0]
    if len(bad):
        ds[bad[0]:] = 0
    return np.concatenate([[0.0], ds])


def _synth():
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                        "mwr_fast_forward_operators_and_lbls_b200", "synth.py")
    spec = importlib.util.spec_from_file_location("_synth_for_tests", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_slant_path_without_trigonometry_matches_serial_ray_tracing():
    """The warp-cooperative slant path never forms theta = 2 asin(s): branch test, cos((theta+theta0)/4)
    and cot(theta) come from s by half-/double-angle identities, small sines/arcsines are series.
    Against the oracle's serial loop (NOAA RAYTRAC as restated), incl. both branches and low angles."""
    d = _synth().synth_profiles(24, 100, seed=3)
    worst = 0.0
    for ip in range(24):
        z, p, t, rh = d["z"][ip], d["p"][ip], d["t"][ip], d["rh"][ip]
        e, _ = ps.vapor(t, rh)
        _, _, n = ps.refractivity(p, t, e)
        z0, zr = z[0], z - z[0]
        for ang in (88.9, 60.0, 30.0, 19.2, 11.4, 6.6, 4.2, 2.0, 1.0):
            ref, got = ps.ray_tracing(zr, n, ang, z0), _slant_kernel_algebra(zr, n, ang, z0)
            assert ((ref == 0) == (got == 0)).all()
            m = ref != 0
            worst = max(worst, float(np.max(np.abs(got[m] - ref[m]) / ref[m])))
    assert worst < 1e-12, worst


def test_slant_path_ducting_zeroes_the_rest_of_the_path():
    """A strong refractivity drop makes argth <= 0 at some level: upstream returns the path so far."""
    d = _synth().synth_profiles(1, 60, seed=5)
    z, p, t, rh = d["z"][0], d["p"][0], d["t"][0], d["rh"][0]
    e, _ = ps.vapor(t, rh)
    _, _, n = ps.refractivity(p, t, e)
    n = n.copy()
    n[8:] = 1.0 + (n[8:] - 1.0) * 1e-3              # elevated duct 0.4 km above the antenna
    ref, got = ps.ray_tracing(z - z[0], n, 0.5, z[0]), _slant_kernel_algebra(z - z[0], n, 0.5, z[0])
    assert (ref[1:] == 0).any() and (ref[1:] != 0).any() and ((ref == 0) == (got == 0)).all()
    m = ref != 0
    assert np.max(np.abs(got[m] - ref[m]) / ref[m]) < 1e-12


def _clog_ratio(a, b):
    nr, ni = a.real * b.real + a.imag * b.imag, a.imag * b.real - a.real * b.imag
    return complex(0.5 * np.log((a.real ** 2 + a.imag ** 2) / (b.real ** 2 + b.imag ** 2)), np.arctan2(ni, nr))


def _liq_unit(model, tk, f):
    """csrc/mwrt_cloud.cuh liq_abs_unit in real arithmetic."""
    if model == 0:
        th1 = 1 - 300 / tk
        eps0 = 77.66 - 103.3 * th1
        eps1, eps2 = 0.0671 * eps0, 3.52
        fp = (316 * th1 + 146.4) * th1 + 20.2
        x1, x2 = f / fp, f / (39.8 * fp)
        a, b = (eps0 - eps1) / (1 + x1 * x1), (eps1 - eps2) / (1 + x2 * x2)
        er, ei = a + b + eps2, -(a * x1 + b * x2)
    else:
        tc, lnth = tk - 273.15, np.log(300 / tk)
        k0 = (-43.7527 * np.exp(0.05 * lnth) + 299.504 * np.exp(1.47 * lnth) - 399.364 * np.exp(2.11 * lnth)
              + 221.327 * np.exp(2.31 * lnth))
        delta, sd = 80.69715 * np.exp(-tc / 226.45), 1164.023 * np.exp(-651.4728 / (tc + 133.07))
        deltab = 4.008724 * np.exp(-tc / 103.05)
        q = 1 / (sd * sd + f * f)
        kr, ki = k0 - delta * f * f * q, -delta * f * sd * q
        f1 = 10.46012 + 0.1454962 * tc + 0.063267156 * tc * tc + 0.00093786645 * tc ** 3
        cn = _clog_ratio(complex(-4500, 2000), complex(-0.75 * f1, f1))
        icn = 1 / (cn.real ** 2 + cn.imag ** 2)
        lp = _clog_ratio(complex(4500, f - 2000), complex(0.75 * f1, f - f1))
        lj = _clog_ratio(complex(4500, f + 2000), complex(0.75 * f1, f + f1))
        chip, chij = lp * cn.conjugate() * (0.5 * deltab * icn), lj * cn * (0.5 * deltab * icn)
        er, ei = kr + (chip.real + chij.real - deltab), ki + (chip.imag + chij.imag)
    return -0.06286 * (3 * ei / ((er + 2) ** 2 + ei * ei)) * f


def test_cloud_liquid_in_real_arithmetic():
    """Im((eps-1)/(eps+2)) = 3 Im eps/|eps+2|^2; z/(sd+z) and log(a/b) without complex division."""
    from oracle import model_tables
    worst = 0.0
    for name, model in (("R98", 0), ("R24", 1)):
        mdl = model_tables.load(name)
        assert (mdl["liq"]["model"] == "liebe91") == (model == 0)
        for tk in np.linspace(233.0, 310.0, 23):
            for f in (1.0, 10.0, 22.24, 31.4, 58.0, 90.0, 183.31, 340.0, 1000.0):
                ref = ps.liquid_water_absorption(mdl, 1.0, f, tk)
                worst = max(worst, abs(_liq_unit(model, tk, f) - ref) / abs(ref))
    assert worst < 1e-12, worst


def test_fp32_pair_form_keeps_the_cancellation_in_fp64():
    """fp32 line-shape mode: vv = (f^2 - fl^2) - (c3 - fl^2) with the first difference rounded ONCE from
    fp64; the pair fraction in fp32 then stays within ~1e-6 of the fp64 value even 5 MHz from a line
    centre, where f^2 - c3 formed in fp32 would lose every digit."""
    n = 20000
    fl = rng.uniform(20, 900, n)
    w = 10 ** rng.uniform(-3, 0.5, n)
    shift = rng.uniform(-0.02, 0.02, n)
    sp = rng.uniform(0.1, 5, n)
    near = rng.uniform(0, 1, n) < 0.5
    f = np.where(near, fl + rng.uniform(-0.5, 0.5, n) * np.maximum(w, 5e-3), rng.uniform(1, 1000, n))
    fc = fl + shift
    fc2, wsq = fc * fc, w * w
    c3, c4, n1 = fc2 - wsq, 4 * fc2 * wsq, 2 * sp * w
    n0 = n1 * 2 * fc2
    vv = f * f - c3
    ref = (n1 * vv + n0) / (vv * vv + c4)
    f32 = np.float32
    a32 = ((f - fl) * (f + fl)).astype(f32)
    b32 = (shift * (2 * fl + shift) - wsq).astype(f32)
    v = a32 - b32
    got = (n1.astype(f32) * v + n0.astype(f32)) / (v * v + c4.astype(f32))
    # where the resolved detuning is not itself below fp32 resolution of the line width
    ok = np.abs(vv) > 1e-4 * np.sqrt(c4)
    err = np.abs(got.astype(np.float64) - ref) / np.abs(ref)
    assert np.max(err[ok]) < 3e-5, np.max(err[ok])
    assert np.median(err) < 3e-7


# ---- far lines on a coarse Chebyshev grid (kernel 1, kNQ = 6) -----------------------------------------
_CHEB6 = np.cos((2 * np.arange(6) + 1) * np.pi / 12.0)
_FAR_RHO = 34.1


def _fast_forms(model, p, tk, e):
    """(c3, c4, n1, n0) of every H2O and O2 line at the level-points (p, tk, e): the pair-loop form of
    csrc/mwrt_absorption.cu (H2OLine / O2Line), in plain NumPy."""
    import json
    import os
    d = json.load(open(os.path.join(os.path.dirname(__file__), "..", "mwr_fast_forward_operators_and_lbls_b200",
                                    "data", "models", model + ".json")))
    h, o = d["h2o"], d["o2"]
    rv = (0.01 * 8.314510) / 18.01528
    rho = e / (rv * tk)
    pvap = rho * tk / h["pvap_div"]
    pda = p - pvap
    ti = h["reftline"] / tk
    tiln = np.log(ti)
    hl = []
    for i in range(len(h["fl"])):
        w = h["w0"][i] * pda * ti ** h["x"][i] + h["w0s"][i] * pvap * ti ** h["xs"][i]
        shift = (h["sr"][i] * w + h["sh"][i] * pda * (1 - h["aair"][i] * tiln) * ti ** h["xh"][i]
                 + h["shs"][i] * pvap * (1 - h["aself"][i] * tiln) * ti ** h["xhs"][i])
        sp = h["s1"][i] * ti ** 2.5 * np.exp(h["b2"][i] * (1 - ti)) / h["fl"][i] ** 2
        fc2, wsq = (h["fl"][i] + shift) ** 2, w * w
        hl.append((h["fl"][i], fc2 - wsq, 4 * fc2 * wsq, 2 * sp * w, 4 * sp * w * fc2))
    th = 300.0 / tk
    th1 = th - 1
    preswv = rho * tk / o["pvap_div"]
    den = 0.001 * ((p - preswv) * th ** o["x"] + o["h2o_broad"] * preswv * th)
    pe2 = den * den
    ol = []
    for k in range(len(o["f"])):
        df = o["w300"][k] * den
        y = den * (o["y0"][k] + o["y1"][k] * th1)
        g = 1 + pe2 * (o["g0"][k] + o["g1"][k] * th1)
        fcen = o["f"][k] + pe2 * (o["dnu0"][k] + o["dnu1"][k] * th1)
        sp = o["s300"][k] * np.exp(-o["be"][k] * th1) / o["f"][k] ** 2
        a_s, pp, fc2, df2 = sp * df * g, sp * y * fcen, fcen ** 2, df * df
        shmax = 4.0 * (abs(o["dnu0"][k]) + abs(o["dnu1"][k])) + 1e-6
        ol.append((o["f"][k], fc2 - df2, 4 * fc2 * df2, 2 * (a_s + pp), 4 * (a_s * fc2 - pp * df2), shmax))
    return hl, ol


def test_far_lines_on_a_chebyshev_grid_match_the_full_sum():
    """Kernel 1 sums the lines that are far from a frequency block (the per-block test of k_abs_blocks:
    |f_line^2 - u0| - margins >= 34.1 half-widths of the block in u = f^2) at 6 Chebyshev nodes and
    interpolates.  Against the full evaluation the line sum must not move by more than 1e-10 (measured:
    3e-11; the parity gate on absorption is 1e-9) -- for the
    HATPRO block, for 28-wide sub-band blocks and for 32-wide blocks of the dense spectrum."""
    from mwr_fast_forward_operators_and_lbls_b200 import synth
    from oracle import pyrtlib_vec as pv
    d = synth.synth_profiles(40, 100, seed=5)
    p, tk = d["p"].ravel(), d["t"].ravel()
    e = pv.vapor(d["t"], d["rh"])[0].ravel()
    hl, ol = _fast_forms("R24", p, tk, e)
    bands = np.asarray(synth.band_frequencies(50)).ravel()
    dense = np.asarray(synth.dense_spectrum())
    blocks = [np.asarray(synth.HATPRO_FRQ)] + [bands[b * 28:(b + 1) * 28] for b in (0, 6, 13, 20, 24)] + \
             [dense[b * 32:(b + 1) * 32] for b in (0, 7, 50, 100, 110, 120)]
    worst, nfar_total = 0.0, 0
    for frq in blocks:
        u = frq ** 2
        u0 = 0.5 * (u.min() + u.max())
        hh = max(0.5 * (u.max() - u.min()), 1e-3 * u0)
        xn = u0 + hh * _CHEB6
        tt = (u - u0) / hh
        M = np.ones((len(u), 6))
        for n in range(6):
            for q in range(6):
                if q != n:
                    M[:, n] *= (tt - _CHEB6[q]) / (_CHEB6[n] - _CHEB6[q])
        for lines, is_h2o in ((hl, True), (ol, False)):
            tot, err = 0.0, 0.0
            for ln in lines:
                fl, c3, c4, n1, n0 = ln[:5]
                if is_h2o and not (fl + frq.max() < 749.0):      # only all-in lines take the pair form
                    continue
                margin = 2.0 * fl * (1.0 if is_h2o else ln[5]) + 37.0
                g = lambda x: (n1[:, None] * (x[None, :] - c3[:, None]) + n0[:, None]) / \
                    ((x[None, :] - c3[:, None]) ** 2 + c4[:, None])
                full = g(u)
                tot = tot + full
                if abs(fl * fl - u0) - margin >= _FAR_RHO * hh:
                    nfar_total += 1
                    err = err + (g(xn) @ M.T - full)
            worst = max(worst, float(np.max(np.abs(err) / np.abs(tot))))
    assert nfar_total > 300          # the criterion is not vacuous: most lines of the narrow blocks are far
    assert worst < 1e-10, worst
