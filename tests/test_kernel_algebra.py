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
