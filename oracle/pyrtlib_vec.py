"""Vectorised NumPy restatement of the `TbCloudRTE` path, batched over profiles.
TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED (no PyRTlib goldens exist).

Written independently of `pyrtlib_scalar.py` (array expressions + masks instead of scalar branches);
`tests/test_oracle.py` requires the two to agree to 1e-12 relative.  Used as the checker for the
CUDA path at sizes of 10^2..10^3 profiles, where the scalar form would take minutes.

Same semantics as the reference call pattern `PyRTlib_processing.py:123-127`:
inputs ground-first, z[km] p[mb] t[K] rh[fraction], frq[GHz], angles[deg elevation].
"""
import numpy as np

from . import model_tables
from .pyrtlib_scalar import (PLANCK_H, BOLTZMANN_K, LIGHT_C, T_COSMIC, EARTH_RADIUS, R_WATVAP,
                             SQRT_PI, EXPMAX, _HUI_A, _HUI_B)

_DB2NP = np.log(10.0) * 0.1
_RVAP = (0.01 * 8.314510) / 18.01528


def dcerror_vec(x, y):
    a, b = _HUI_A, _HUI_B
    zh = np.abs(y) - 1j * x
    asum = (((((a[6] * zh + a[5]) * zh + a[4]) * zh + a[3]) * zh + a[2]) * zh + a[1]) * zh + a[0]
    bsum = ((((((zh + b[6]) * zh + b[5]) * zh + b[4]) * zh + b[3]) * zh + b[2]) * zh + b[1]) * zh + b[0]
    w = asum / bsum
    neg = y < 0
    if np.any(neg):
        w = np.where(neg, 2.0 * np.exp(-(x + 1j * y) ** 2) - np.conj(w), w)
    return w


def vapor(tk, rh):
    y = 373.16 / tk
    es = (-7.90298 * (y - 1.0) + 5.02808 * np.log10(y)
          - 1.3816e-07 * (10 ** (11.344 * (1.0 - (1.0 / y))) - 1.0)
          + 0.0081328 * (10 ** (-3.49149 * (y - 1.0)) - 1.0) + np.log10(1013.246))
    es = 10.0 ** es
    e = rh * es
    rho = e / ((R_WATVAP * 1e-05) * tk)
    return e, rho


def h2o_abs(mdl, p, tk, e, f):
    """H2O absorption [Np/km] for arrays p,tk,e (any shape) at scalar frequency f."""
    h = mdl["h2o"]
    ekpa = e / 10.0
    vx = 300.0 / tk
    t = 300.0 / vx
    pdrykpa = p / 10.0 - ekpa
    pp = (pdrykpa + ekpa) * 10.0
    rho = ekpa * 10.0 / (_RVAP * t)
    pvap = (rho * t) / h["pvap_div"]
    pda = pp - pvap
    den = h["den_const"] * rho
    ti = h["reftcon"] / t
    con = (h["cf"] * pda * ti ** h["xcf"] + h["cs"] * pvap * ti ** h["xcs"]) * pvap * f * f
    ti = h["reftline"] / t
    tiln = np.log(ti)
    ti2 = np.exp(2.5 * tiln)
    summ = np.zeros_like(t)
    for i in range(len(h["fl"])):
        width0 = h["w0"][i] * pda * ti ** h["x"][i] + h["w0s"][i] * pvap * ti ** h["xs"][i]
        if h["sd"]:
            width2 = h["w2"][i] * pda + h["w2s"][i] * pvap
            delta2 = h["d2"][i] * pda + h["d2s"][i] * pvap
        else:
            width2 = np.zeros_like(t)
            delta2 = np.zeros_like(t)
        shiftf = h["sh"][i] * pda * (1.0 - h["aair"][i] * tiln) * ti ** h["xh"][i]
        shifts = h["shs"][i] * pvap * (1.0 - h["aself"][i] * tiln) * ti ** h["xhs"][i]
        shift = h["sr"][i] * width0 + shiftf + shifts
        wsq = width0 ** 2
        s = h["s1"][i] * ti2 * np.exp(h["b2"][i] * (1.0 - ti))
        df0 = f - h["fl"][i] - shift
        df1 = f + h["fl"][i] + shift
        base = width0 / (562500.0 + wsq)
        lor0 = np.where(np.abs(df0) < 750.0, width0 / (df0 ** 2 + wsq) - base, 0.0)
        lor1 = np.where(np.abs(df1) < 750.0, width0 / (df1 ** 2 + wsq) - base, 0.0)
        use_sd = (width2 > 0.0) & (np.abs(df0) < 10.0 * width0)
        if np.any(use_sd):
            w2s = np.where(use_sd, width2, 1.0)
            cden = w2s - 1j * delta2
            xc = ((width0 - 1.5 * w2s) + 1j * (df0 + 1.5 * delta2)) / cden
            xrt = np.sqrt(xc)
            pxw = SQRT_PI * xrt * dcerror_vec(-xrt.imag, xrt.real)
            sd = 2.0 * (1.0 - pxw) / cden
            lor0 = np.where(use_sd, sd.real - base, lor0)
        res = lor0 + lor1
        summ = summ + s * res * (f / h["fl"][i]) ** 2
    habs = h["line_const"] * den * summ + con
    # upstream splits into npp/ncpp (divide by db2np*factor) and clearsky_absorption multiplies back
    factor = 0.182 * f
    npp = (h["line_const"] * den * summ) / _DB2NP / factor
    ncpp = con / _DB2NP / factor
    habs = (factor * (npp + ncpp)) * _DB2NP
    return np.where(rho <= 0.0, 0.0, habs)


def o2_abs(mdl, p, tk, e, f):
    o = mdl["o2"]
    ekpa = e / 10.0
    vx = 300.0 / tk
    temp = 300.0 / vx
    pdrykpa = p / 10.0 - ekpa
    pres = (pdrykpa + ekpa) * 10.0
    vapden = ekpa * 10.0 / (_RVAP * temp)
    th = 300.0 / temp
    th1 = th - 1.0
    b = th ** o["x"]
    preswv = vapden * temp / o["pvap_div"]
    presda = pres - preswv
    den = 0.001 * (presda * b + o["h2o_broad"] * preswv * th)
    dens = 0.001 * (presda + o["h2o_broad"] * preswv) * th
    dfnr = o["wb300"] * den
    pe2 = den * den
    summ = o["nonres"] * f * f * dfnr / (th * (f * f + dfnr * dfnr))
    for k in range(len(o["f"])):
        df = o["w300"][k] * den
        if o["dens118"] and k == 0:
            df = o["w300"][0] * dens
        if o["mixing"] == 1:
            y = 0.001 * pres * b * (o["y0"][k] + o["y1"][k] * th1)
            gfac = np.ones_like(th)
            fcen = np.full_like(th, o["f"][k])
        else:
            y = den * (o["y0"][k] + o["y1"][k] * th1)
            gfac = 1.0 + pe2 * (o["g0"][k] + o["g1"][k] * th1)
            fcen = o["f"][k] + pe2 * (o["dnu0"][k] + o["dnu1"][k] * th1)
        strr = o["s300"][k] * np.exp(-o["be"][k] * th1)
        sf1 = (df * gfac + (f - fcen) * y) / ((f - fcen) ** 2 + df * df)
        if o["sd118"] and k == 0:
            use_sd = np.abs(f - fcen) < 10.0 * df
            if np.any(use_sd):
                width2 = 0.076 * df
                xc = ((df - 1.5 * width2) + 1j * (f - fcen)) / width2
                xrt = np.sqrt(xc)
                pxw = SQRT_PI * xrt * dcerror_vec(-xrt.imag, xrt.real)
                asd = (1.0 + 1j * y) * 2.0 * (1.0 - pxw) / width2
                sf1 = np.where(use_sd, asd.real, sf1)
        sf2 = (df * gfac - (f + fcen) * y) / ((f + fcen) ** 2 + df * df)
        summ = summ + strr * (sf1 + sf2) * (f / o["f"][k]) ** 2
    o2abs = o["abs_num"] * summ * presda * th ** 3 / o["abs_den"]
    o2abs = o["iso"] * np.maximum(o2abs, 0.0)
    factor = 0.182 * f
    npp = o2abs / _DB2NP / factor
    return (factor * npp) * _DB2NP


def n2_abs(mdl, p, tk, e, f):
    n2 = mdl["n2"]
    pdry = (p / 10.0 - e / 10.0) * 10.0
    th = 300.0 / tk
    fdepen = 0.5 + 0.5 / (1.0 + (f / 450.0) ** 2) if n2["fdep"] else 1.0
    bf = n2["l"] * fdepen * pdry * pdry * f * f * th ** n2["m"]
    return n2["n"] * bf


def dilec12_vec(f, tk):
    tc = tk - 273.15
    z = 1j * f
    theta = 300.0 / tk
    kappa = (-43.7527 * theta ** 0.05 + 299.504 * theta ** 1.47 - 399.364 * theta ** 2.11
             + 221.327 * theta ** 2.31)
    delta = 80.69715 * np.exp(-tc / 226.45)
    sd = 1164.023 * np.exp(-651.4728 / (tc + 133.07))
    kappa = kappa - delta * z / (sd + z)
    delta = 4.008724 * np.exp(-tc / 103.05)
    hdelta = delta / 2.0
    f1 = 10.46012 + 0.1454962 * tc + 0.063267156 * tc ** 2 + 0.00093786645 * tc ** 3
    z1 = (-0.75 + 1j) * f1
    z2 = -4500.0 + 2000.0j
    cnorm = np.log(z2 / z1)
    chip = (hdelta * np.log((z - z2) / (z - z1))) / cnorm
    chij = (hdelta * np.log((z - np.conj(z2)) / (z - np.conj(z1)))) / np.conj(cnorm)
    return kappa + (chip + chij - delta)


def liq_abs(mdl, tk, denl, f):
    if mdl["liq"]["model"] == "liebe91":
        theta1 = 1.0 - 300.0 / tk
        eps0 = 77.66 - 103.3 * theta1
        eps1 = 0.0671 * eps0
        eps2 = 3.52
        fp = (316.0 * theta1 + 146.4) * theta1 + 20.2
        fs = 39.8 * fp
        eps = (eps0 - eps1) / (1.0 + 1j * (f / fp)) + (eps1 - eps2) / (1.0 + 1j * (f / fs)) + eps2
    else:
        eps = dilec12_vec(f, tk)
    re = (eps - 1.0) / (eps + 2.0)
    return np.where(denl > 0.0, -0.06286 * re.imag * f * denl, 0.0)


def ice_abs(deni, f):
    wave = LIGHT_C / (f * 1e9)
    return np.where(deni > 0.0, ((8.18645 / wave) * deni * 0.000959553) * _DB2NP, 0.0)


def absorption(mdl, p, tk, e, frq, denliq=None, denice=None):
    """awet, adry, aliq, aice as (P, F, L) arrays."""
    mdl = model_tables.load(mdl)
    P, L = p.shape
    F = len(frq)
    awet = np.empty((P, F, L))
    adry = np.empty((P, F, L))
    aliq = np.zeros((P, F, L))
    aice = np.zeros((P, F, L))
    for j, f in enumerate(frq):
        awet[:, j, :] = h2o_abs(mdl, p, tk, e, f)
        adry[:, j, :] = o2_abs(mdl, p, tk, e, f) + n2_abs(mdl, p, tk, e, f)
        if denliq is not None:
            aliq[:, j, :] = liq_abs(mdl, tk, denliq, f)
        if denice is not None:
            aice[:, j, :] = ice_abs(denice, f)
    return awet, adry, aliq, aice


def layer_mean(x, zeroflg):
    """Per-layer representative value of x (..., L) -> (..., L) with [...,0] = 0
    (`exponential_integration` without the ds factor).  Negative x -> NaN (upstream raises)."""
    x0 = x[..., :-1]
    x1 = x[..., 1:]
    with np.errstate(divide="ignore", invalid="ignore"):
        lm = (x1 - x0) / np.log(x1 / x0)
    anyzero = (x0 == 0.0) | (x1 == 0.0)
    lay = np.where(anyzero, (x1 + x0) * 0.5 if zeroflg else 0.0, lm)
    lay = np.where(np.abs(x1 - x0) < 1e-09, x1, lay)
    lay = np.where((x0 < 0.0) | (x1 < 0.0), np.nan, lay)
    out = np.zeros_like(x)
    out[..., 1:] = lay
    return out


def refractivity(p, tk, e):
    pa = p - e
    tc = tk - 273.16
    tk2 = tk * tk
    tc2 = tc * tc
    rza = 1.0 + pa * (5.79e-07 * (1.0 + 0.52 / tk) - 0.00094611 * tc / tk2)
    rzw = 1.0 + 1650.0 * (e / (tk * tk2)) * (1.0 - 0.01317 * tc + 0.000175 * tc2
                                              + 1.44e-06 * (tc2 * tc))
    wetn = (64.79 * (e / tk) + 377600.0 * (e / tk2)) * rzw
    dryn = 77.6036 * (pa / tk) * rza
    return 1.0 + (dryn + wetn) * 1e-06


def slant_path(z, p, tk, e, angles, ray_tracing):
    """ds (P, A, L) [km]; z is absolute height, made relative to z[:,0] here."""
    from .pyrtlib_scalar import ray_tracing as rt_scalar
    P, L = z.shape
    A = len(angles)
    zr = z - z[:, :1]
    ds = np.zeros((P, A, L))
    if not ray_tracing:
        for k, ang in enumerate(angles):
            amass = 1.0 / np.sin(ang * np.pi / 180.0)
            ds[:, k, 1:] = (zr[:, 1:] - zr[:, :-1]) * amass
        return ds
    refindx = refractivity(p, tk, e)
    for i in range(P):
        for k, ang in enumerate(angles):
            ds[i, k, :] = rt_scalar(zr[i], refindx[i], float(ang), float(z[i, 0]))
    return ds


def rte(tk, frq, tau_lay, from_sat=False, es=1.0):
    """Downwelling (or upwelling) layer recursion.  tk (P,L), tau_lay (P,F,A,L) -> dict of (P,F,A)."""
    P, F, A, L = tau_lay.shape
    hvk = (np.asarray(frq) * 1e9) * PLANCK_H / BOLTZMANN_K          # (F,)
    boft = 1.0 / (np.exp(hvk[None, :, None] / tk[:, None, :]) - 1.0)  # (P,F,L)
    tauprof = np.zeros((P, F, A))
    boftatm = np.zeros((P, F, A))
    if not from_sat:
        for i in range(1, L):
            ex = np.exp(-tau_lay[..., i])
            boftlay = (boft[:, :, None, i - 1] + boft[:, :, None, i] * ex) / (1.0 + ex)
            batmlay = boftlay * np.exp(-tauprof) * (1.0 - ex)
            boftatm = boftatm + batmlay
            tauprof = tauprof + tau_lay[..., i]
        bg = 1.0 / (np.exp(hvk / T_COSMIC) - 1.0)
        bg = bg[None, :, None]
    else:
        for i in range(L - 2, -1, -1):
            ex = np.exp(-tau_lay[..., i + 1])
            boftlay = (boft[:, :, None, i + 1] + boft[:, :, None, i] * ex) / (1.0 + ex)
            batmlay = boftlay * np.exp(-tauprof) * (1.0 - ex)
            boftatm = boftatm + batmlay
            tauprof = tauprof + tau_lay[..., i + 1]
        bg = es * boft[:, :, None, 0]
    thin = tauprof < EXPMAX
    with np.errstate(divide="ignore", invalid="ignore"):
        boftotl = np.where(thin, bg * np.exp(-tauprof) + boftatm, boftatm)
        boftmr = np.where(thin, boftatm / (1.0 - np.exp(-tauprof)), boftatm)
        h3 = hvk[None, :, None]

        def bright(b):
            return np.where((b > 0.0) | np.isnan(b), h3 / np.log(1.0 + 1.0 / b), 0.0)
        return dict(tbtotal=bright(boftotl), tbatm=bright(boftatm), tmr=bright(boftmr),
                    tauprof=tauprof)


def simulate(z, p, tk, rh, frq, angles, model, ray_tracing=False, from_sat=False, denliq=None,
             denice=None, return_layers=False):
    """Batched forward model: inputs (P,L) ground-first; returns dict of (P,F,A) arrays
    tbtotal, tbatm, tmr, tauwet, taudry, tauliq, tauice (+ awet/adry/... (P,F,L) if return_layers).
    A profile containing NaN anywhere returns NaN everywhere (reference gate,
    PyRTlib_processing.py:71-79,117-119)."""
    mdl = model_tables.load(model)
    z = np.asarray(z, dtype=np.float64)
    p = np.asarray(p, dtype=np.float64)
    tk = np.asarray(tk, dtype=np.float64)
    rh = np.asarray(rh, dtype=np.float64)
    frq = np.atleast_1d(np.asarray(frq, dtype=np.float64))
    angles = np.atleast_1d(np.asarray(angles, dtype=np.float64))
    e, rho = vapor(tk, rh)
    cloudy = denliq is not None or denice is not None
    if cloudy:
        denliq = np.zeros_like(tk) if denliq is None else np.asarray(denliq, dtype=np.float64)
        denice = np.zeros_like(tk) if denice is None else np.asarray(denice, dtype=np.float64)
    awet, adry, aliq, aice = absorption(mdl, p, tk, e, frq, denliq, denice)
    ds = slant_path(z, p, tk, e, angles, ray_tracing)                 # (P,A,L)
    lw = layer_mean(awet, True)                                       # (P,F,L)
    ld = layer_mean(adry, True)
    ll = layer_mean(aliq, False)
    li = layer_mean(aice, False)
    dsx = ds[:, None, :, :]                                           # (P,1,A,L)
    tw = lw[:, :, None, :] * dsx                                      # (P,F,A,L)
    td = ld[:, :, None, :] * dsx
    tl = ll[:, :, None, :] * dsx
    ti = li[:, :, None, :] * dsx

    def seqsum(x):  # serial sum over layers, the order the scalar form uses
        acc = np.zeros(x.shape[:-1])
        for i in range(1, x.shape[-1]):
            acc = acc + x[..., i]
        return acc
    out = dict(tauwet=seqsum(tw), taudry=seqsum(td), tauliq=seqsum(tl), tauice=seqsum(ti))
    tau_lay = tw + td + ti + tl
    out.update(rte(tk, frq, tau_lay, from_sat=from_sat))
    out.pop("tauprof")
    bad = (np.isnan(z).any(1) | np.isnan(p).any(1) | np.isnan(tk).any(1) | np.isnan(rh).any(1))
    for k in list(out):
        out[k] = np.where(bad[:, None, None], np.nan, out[k])
    if return_layers:
        out.update(awet=awet, adry=adry, aliq=aliq, aice=aice, ds=ds, e=e, rho=rho,
                   tau_layer=tau_lay)
    return out
