"""PyRTlib-shaped scalar restatement of the `TbCloudRTE` path.  TEST INFRASTRUCTURE (see oracle/__init__.py).

PARITY UNPINNED: restated from memory of PyRTlib 1.1.x / Rosenkranz's Fortran; no golden vectors exist
for this path in the reference.

Shape: exactly the call pattern the reference uses --
    rte = TbCloudRTE(z, p, t, rh, frq, ang); rte.init_absmdl(mdl); rte.satellite = False; df = rte.execute()
(`/root/reference/python_src/proc/PyRTlib_processing.py:123-127`) -- implemented the way the upstream
engine does it: Python loops over angles, frequencies, levels and lines, absorption recomputed for
every angle.  It is deliberately slow; it is the CPU arm `bench.py` times.

Upstream files restated (not present under /root/reference; names per SURVEY.md 8(c)):
  pyrtlib/tb_spectrum.py      TbCloudRTE.__init__/init_absmdl/init_cloudy/execute
  pyrtlib/rt_equation.py      RTEquation.vapor/refractivity/ray_tracing/exponential_integration/
                              clearsky_absorption/cloudy_absorption/planck/bright
  pyrtlib/absorption_model.py H2OAbsModel/O2AbsModel/N2AbsModel/LiqAbsModel
  pyrtlib/utils.py            constants, dcerror, dilec12
"""
import cmath
import math

import numpy as np

from . import model_tables

# pyrtlib.utils.constants (values as recalled; SURVEY.md A.10 flags Tc as "check exact constant")
PLANCK_H = 6.6260755e-34      # J s
BOLTZMANN_K = 1.380658e-23    # J/K
LIGHT_C = 299792458.0         # m/s
T_COSMIC = 2.736              # K: constants('Tcosmicbkg'), 2.736 +/- 0.017 K (Janssen, Atmospheric Remote Sensing
                              # by Microwave Radiometry, p. 12) -- the value of the NOAA/Cimini RTE package
                              # PyRTlib was ported from; round 1 carried 2.73 (a ~5e-3 K K-band offset)
EARTH_RADIUS = 6370.949       # km
R_WATVAP = 461.5              # J/(kg K)
SQRT_PI = 1.77245385090551603
EXPMAX = 125.0


# --------------------------------------------------------------------------------------------
# pyrtlib.utils
# --------------------------------------------------------------------------------------------
_HUI_A = (122.607931777104326, 214.382388694706425, 181.928533092181549, 93.155580458138441,
          30.180142196210589, 5.912626209773153, 0.564189583562615)
_HUI_B = (122.607931773875350, 352.730625110963558, 457.334478783897737, 348.703917719495792,
          170.354001821091472, 53.992906912940207, 10.479857114260399)


def dcerror(x, y):
    """Complex error function w(x+iy), Hui/Armstrong/Wray 6th-order rational."""
    a, b = _HUI_A, _HUI_B
    zh = complex(abs(y), -x)
    asum = (((((a[6] * zh + a[5]) * zh + a[4]) * zh + a[3]) * zh + a[2]) * zh + a[1]) * zh + a[0]
    bsum = ((((((zh + b[6]) * zh + b[5]) * zh + b[4]) * zh + b[3]) * zh + b[2]) * zh + b[1]) * zh + b[0]
    w = asum / bsum
    if y >= 0:
        return w
    return 2.0 * cmath.exp(-complex(x, y) ** 2) - w.conjugate()


def dilec12(f, tk):
    """Complex permittivity of liquid water, Rosenkranz (2015)."""
    tc = tk - 273.15
    z = complex(0.0, f)
    theta = 300.0 / tk
    kappa = (-43.7527 * theta ** 0.05 + 299.504 * theta ** 1.47 - 399.364 * theta ** 2.11
             + 221.327 * theta ** 2.31)
    delta = 80.69715 * math.exp(-tc / 226.45)
    sd = 1164.023 * math.exp(-651.4728 / (tc + 133.07))
    kappa = kappa - delta * z / (sd + z)
    delta = 4.008724 * math.exp(-tc / 103.05)
    hdelta = delta / 2.0
    f1 = 10.46012 + 0.1454962 * tc + 0.063267156 * tc ** 2 + 0.00093786645 * tc ** 3
    z1 = complex(-0.75, 1.0) * f1
    z2 = complex(-4500.0, 2000.0)
    cnorm = cmath.log(z2 / z1)
    chip = (hdelta * cmath.log((z - z2) / (z - z1))) / cnorm
    chij = (hdelta * cmath.log((z - z2.conjugate()) / (z - z1.conjugate()))) / cnorm.conjugate()
    dchi = chip + chij - delta
    return kappa + dchi


# --------------------------------------------------------------------------------------------
# pyrtlib.absorption_model
# --------------------------------------------------------------------------------------------
_DB2NP = math.log(10.0) * 0.1
_RVAP = (0.01 * 8.314510) / 18.01528


def h2o_absorption(mdl, pdrykpa, vx, ekpa, frq):
    """Rosenkranz `abh2o` as wrapped by PyRTlib: returns (npp, ncpp)."""
    h = mdl["h2o"]
    factor = 0.182 * frq
    t = 300.0 / vx
    p = (pdrykpa + ekpa) * 10.0
    rho = ekpa * 10.0 / (_RVAP * t)
    f = frq
    if rho <= 0.0:
        return 0.0, 0.0
    pvap = (rho * t) / h["pvap_div"]
    pda = p - pvap
    den = h["den_const"] * rho
    # continuum terms
    ti = h["reftcon"] / t
    con = (h["cf"] * pda * ti ** h["xcf"] + h["cs"] * pvap * ti ** h["xcs"]) * pvap * f * f
    # add resonances
    ti = h["reftline"] / t
    tiln = math.log(ti)
    ti2 = math.exp(2.5 * tiln)
    summ = 0.0
    for i in range(len(h["fl"])):
        width0 = h["w0"][i] * pda * ti ** h["x"][i] + h["w0s"][i] * pvap * ti ** h["xs"][i]
        width2 = 0.0
        delta2 = 0.0
        if h["sd"]:
            width2 = h["w2"][i] * pda + h["w2s"][i] * pvap
            delta2 = h["d2"][i] * pda + h["d2s"][i] * pvap
        shiftf = h["sh"][i] * pda * (1.0 - h["aair"][i] * tiln) * ti ** h["xh"][i]
        shifts = h["shs"][i] * pvap * (1.0 - h["aself"][i] * tiln) * ti ** h["xhs"][i]
        shift = h["sr"][i] * width0 + shiftf + shifts
        wsq = width0 ** 2
        s = h["s1"][i] * ti2 * math.exp(h["b2"][i] * (1.0 - ti))
        df = (f - h["fl"][i] - shift, f + h["fl"][i] + shift)
        base = width0 / (562500.0 + wsq)
        res = 0.0
        for j in range(2):
            if width2 > 0.0 and j == 0 and abs(df[j]) < 10.0 * width0:
                # speed-dependent resonant shape factor
                xc = complex(width0 - 1.5 * width2, df[j] + 1.5 * delta2) / complex(width2, -delta2)
                xrt = cmath.sqrt(xc)
                pxw = SQRT_PI * xrt * dcerror(-xrt.imag, xrt.real)
                sd = 2.0 * (1.0 - pxw) / complex(width2, -delta2)
                res += sd.real - base
            elif abs(df[j]) < 750.0:
                res += width0 / (df[j] ** 2 + wsq) - base
        summ += s * res * (f / h["fl"][i]) ** 2
    npp = (h["line_const"] * den * summ) / _DB2NP / factor
    ncpp = con / _DB2NP / factor
    return npp, ncpp


def o2_absorption(mdl, pdrykpa, vx, ekpa, frq):
    """Rosenkranz `o2abs` as wrapped by PyRTlib: returns (npp, ncpp)."""
    o = mdl["o2"]
    factor = 0.182 * frq
    temp = 300.0 / vx
    pres = (pdrykpa + ekpa) * 10.0
    vapden = ekpa * 10.0 / (_RVAP * temp)
    freq = frq
    th = 300.0 / temp
    th1 = th - 1.0
    b = th ** o["x"]
    preswv = vapden * temp / o["pvap_div"]
    presda = pres - preswv
    den = 0.001 * (presda * b + o["h2o_broad"] * preswv * th)
    dens = 0.001 * (presda + o["h2o_broad"] * preswv) * th
    dfnr = o["wb300"] * den
    pe2 = den * den
    summ = o["nonres"] * freq * freq * dfnr / (th * (freq * freq + dfnr * dfnr))
    for k in range(len(o["f"])):
        df = o["w300"][k] * den
        if o["dens118"] and k == 0:
            df = o["w300"][0] * dens
        if o["mixing"] == 1:
            y = 0.001 * pres * b * (o["y0"][k] + o["y1"][k] * th1)
            gfac = 1.0
            fcen = o["f"][k]
        else:
            y = den * (o["y0"][k] + o["y1"][k] * th1)
            gfac = 1.0 + pe2 * (o["g0"][k] + o["g1"][k] * th1)
            fcen = o["f"][k] + pe2 * (o["dnu0"][k] + o["dnu1"][k] * th1)
        strr = o["s300"][k] * math.exp(-o["be"][k] * th1)
        if o["sd118"] and k == 0 and abs(freq - fcen) < 10.0 * df:
            width2 = 0.076 * df
            xc = complex(df - 1.5 * width2, freq - fcen) / width2
            xrt = cmath.sqrt(xc)
            pxw = SQRT_PI * xrt * dcerror(-xrt.imag, xrt.real)
            asd = complex(1.0, y) * 2.0 * (1.0 - pxw) / width2
            sf1 = asd.real
        else:
            sf1 = (df * gfac + (freq - fcen) * y) / ((freq - fcen) ** 2 + df * df)
        sf2 = (df * gfac - (freq + fcen) * y) / ((freq + fcen) ** 2 + df * df)
        summ += strr * (sf1 + sf2) * (freq / o["f"][k]) ** 2
    o2abs = o["abs_num"] * summ * presda * th ** 3 / o["abs_den"]
    o2abs = o["iso"] * max(o2abs, 0.0)
    npp = o2abs / _DB2NP / factor
    ncpp = 0.0  # only npp+ncpp is used downstream (clearsky_absorption)
    return npp, ncpp


def n2_absorption(mdl, t, p, f):
    """Collision-induced N2 absorption [Np/km]; p = dry pressure [mb]."""
    n2 = mdl["n2"]
    th = 300.0 / t
    fdepen = 0.5 + 0.5 / (1.0 + (f / 450.0) ** 2)
    if not n2["fdep"]:
        fdepen = 1.0
    bf = n2["l"] * fdepen * p * p * f * f * th ** n2["m"]
    return n2["n"] * bf


def liquid_water_absorption(mdl, water, freq, temp):
    """Rayleigh liquid-water absorption [Np/km]; water in g/m3."""
    if water <= 0.0:
        return 0.0
    if mdl["liq"]["model"] == "liebe91":
        theta1 = 1.0 - 300.0 / temp
        eps0 = 77.66 - 103.3 * theta1
        eps1 = 0.0671 * eps0
        eps2 = 3.52
        fp = (316.0 * theta1 + 146.4) * theta1 + 20.2
        fs = 39.8 * fp
        eps = ((eps0 - eps1) / complex(1.0, freq / fp) + (eps1 - eps2) / complex(1.0, freq / fs)
               + eps2)
    else:
        eps = dilec12(freq, temp)
    re = (eps - 1.0) / (eps + 2.0)
    return -0.06286 * re.imag * freq * water


# --------------------------------------------------------------------------------------------
# pyrtlib.rt_equation.RTEquation
# --------------------------------------------------------------------------------------------
def vapor(tk, rh):
    """Goff-Gratch saturation pressure over water; e = rh*es [mb], rho [g/m3]."""
    rvap = R_WATVAP * 1e-05
    tk = np.asarray(tk, dtype=np.float64)
    rh = np.asarray(rh, dtype=np.float64)
    e = np.zeros(tk.shape)
    rho = np.zeros(tk.shape)
    for i in range(len(tk)):
        y = 373.16 / tk[i]
        es = (-7.90298 * (y - 1.0) + 5.02808 * math.log10(y)
              - 1.3816e-07 * (10 ** (11.344 * (1.0 - (1.0 / y))) - 1.0)
              + 0.0081328 * (10 ** (-3.49149 * (y - 1.0)) - 1.0) + math.log10(1013.246))
        es = 10.0 ** es
        e[i] = rh[i] * es
        rho[i] = e[i] / (rvap * tk[i])
    return e, rho


def refractivity(p, tk, e):
    """NOAA REFRACT: returns dryn, wetn, refindx."""
    nl = len(p)
    dryn = np.zeros(nl)
    wetn = np.zeros(nl)
    refindx = np.zeros(nl)
    for i in range(nl):
        pa = p[i] - e[i]
        tc = tk[i] - 273.16
        tk2 = tk[i] * tk[i]
        tc2 = tc * tc
        rza = 1.0 + pa * (5.79e-07 * (1.0 + 0.52 / tk[i]) - 0.00094611 * tc / tk2)
        rzw = 1.0 + 1650.0 * (e[i] / (tk[i] * tk2)) * (1.0 - 0.01317 * tc + 0.000175 * tc2
                                                        + 1.44e-06 * (tc2 * tc))
        wetn[i] = (64.79 * (e[i] / tk[i]) + 377600.0 * (e[i] / tk2)) * rzw
        dryn[i] = 77.6036 * (pa / tk[i]) * rza
        refindx[i] = 1.0 + (dryn[i] + wetn[i]) * 1e-06
    return dryn, wetn, refindx


def ray_tracing(z, refindx, angle, z0):
    """NOAA RAYTRAC: refracted slant path [km] per layer, ds[0] = 0.
    z is height above the antenna [km], z0 the antenna height [km], angle = elevation [deg].
    On ducting the already-computed part of ds is returned (rest zero), like upstream."""
    deg2rad = math.pi / 180.0
    re = EARTH_RADIUS
    nl = len(z)
    ds = np.zeros(nl)
    for i in range(nl):
        if refindx[i] < 1.0:
            return ds
    if (89.0 <= angle <= 91.0) or angle == -90.0:
        ds[0] = 0.0
        for i in range(1, nl):
            ds[i] = z[i] - z[i - 1]
        return ds
    theta0 = angle * deg2rad
    rs = re + z[0] + z0
    costh0 = math.cos(theta0)
    sina = math.sin(theta0 * 0.5)
    a0 = 2.0 * sina ** 2
    ds[0] = 0.0
    phil = 0.0
    taul = 0.0
    rl = re + z[0] + z0
    tanthl = math.tan(theta0)
    for i in range(1, nl):
        r = re + z[i] + z0
        if refindx[i] == refindx[i - 1] or refindx[i] == 1.0 or refindx[i - 1] == 1.0:
            refbar = (refindx[i] + refindx[i - 1]) * 0.5
        else:
            refbar = 1.0 + (refindx[i - 1] - refindx[i]) / (
                math.log((refindx[i - 1] - 1.0) / (refindx[i] - 1.0)))
        argdth = z[i] / rs - (refindx[0] - refindx[i]) * costh0 / refindx[i]
        argth = 0.5 * (a0 + argdth) / r
        if argth <= 0:
            return ds  # ducting
        sint = math.sqrt(r * argth)
        theta = 2.0 * math.asin(sint)
        if (theta - 2.0 * theta0) <= 0.0:
            dendth = 2.0 * (sint + sina) * math.cos((theta + theta0) * 0.25)
            sind4 = (0.5 * argdth - z[i] * argth) / dendth
            dtheta = 4.0 * math.asin(sind4)
            theta = theta0 + dtheta
        else:
            dtheta = theta - theta0
        tanth = math.tan(theta)
        cthbar = ((1.0 / tanth) + (1.0 / tanthl)) * 0.5
        dtau = cthbar * (refindx[i - 1] - refindx[i]) / refbar
        tau = taul + dtau
        phi = dtheta + tau
        ds[i] = math.sqrt((z[i] - z[i - 1]) ** 2
                          + 4.0 * r * rl * ((math.sin((phi - phil) * 0.5)) ** 2))
        if dtau != 0.0:
            dtaua = abs(tau - taul)
            ds[i] = ds[i] * (dtaua / (2.0 * math.sin(dtaua * 0.5)))
        phil = phi
        taul = tau
        rl = r
        tanthl = tanth
    return ds


def exponential_integration(zeroflg, x, ds, ibeg, iend, factor):
    """Layer integrals assuming exponential variation within a layer."""
    sxds = 0.0
    xds = np.zeros(len(ds))
    for i in range(ibeg + 1, iend):
        if x[i - 1] < 0.0 or x[i] < 0.0:
            raise ValueError("Error encountered in exponential_integration")
        elif abs(x[i] - x[i - 1]) < 1e-09:
            xlayer = x[i]
        elif x[i - 1] == 0.0 or x[i] == 0.0:
            if zeroflg:
                xlayer = (x[i] + x[i - 1]) * 0.5
            else:
                xlayer = 0.0
        else:
            xlayer = (x[i] - x[i - 1]) / math.log(x[i] / x[i - 1])
        xds[i] = xlayer * ds[i]
        sxds = sxds + xds[i]
    sxds = sxds * factor
    return sxds, xds


def clearsky_absorption(mdl, p, tk, e, frq):
    """awet, adry [Np/km] at every level for one frequency."""
    nl = len(p)
    awet = np.zeros(nl)
    adry = np.zeros(nl)
    factor = 0.182 * frq
    db2np = math.log(10.0) * 0.1
    for i in range(nl):
        v = 300.0 / tk[i]
        ekpa = e[i] / 10.0
        pdrykpa = p[i] / 10.0 - ekpa
        npp, ncpp = h2o_absorption(mdl, pdrykpa, v, ekpa, frq)
        awet[i] = (factor * (npp + ncpp)) * db2np
        npp, ncpp = o2_absorption(mdl, pdrykpa, v, ekpa, frq)
        ao2 = (factor * (npp + ncpp)) * db2np
        an2 = n2_absorption(mdl, tk[i], pdrykpa * 10.0, frq)
        adry[i] = ao2 + an2
    return awet, adry


def cloudy_absorption(mdl, tk, denl, deni, frq):
    """aliq, aice [Np/km] at every level."""
    nl = len(tk)
    wave = LIGHT_C / (frq * 1e9)
    db2np = math.log(10.0) * 0.1
    aliq = np.zeros(nl)
    aice = np.zeros(nl)
    for i in range(nl):
        if denl[i] > 0:
            aliq[i] = liquid_water_absorption(mdl, denl[i], frq, tk[i])
        if deni[i] > 0:
            aice[i] = (8.18645 / wave) * deni[i] * 0.000959553
            aice[i] = aice[i] * db2np
    return aliq, aice


def tk2b_mod(hvk, t):
    return 1.0 / (math.exp(hvk / t) - 1.0)


def bright(hvk, boft):
    """Inverse (modified) Planck function -> brightness temperature [K]."""
    if boft > 0.0 or boft != boft:
        return hvk / math.log(1.0 + (1.0 / boft))
    return 0.0


def planck(frq, tk, taulay, from_sat=False, es=1.0):
    """Layer recursion of the RTE in (modified) Planck radiance.
    from_sat=False: downwelling seen from the ground, cosmic background at the top.
    from_sat=True : upwelling seen from above; background = es*B(tk[0]) (surface skin = lowest level)."""
    fhz = frq * 1e9
    hvk = fhz * PLANCK_H / BOLTZMANN_K
    nl = len(tk)
    tauprof = np.zeros(nl)
    boftatm = np.zeros(nl)
    boft = np.zeros(nl)
    if from_sat:
        boft[nl - 1] = tk2b_mod(hvk, tk[nl - 1])
        for i in range(nl - 2, -1, -1):
            boft[i] = tk2b_mod(hvk, tk[i])
            ex = math.exp(-taulay[i + 1])
            boftlay = (boft[i + 1] + boft[i] * ex) / (1.0 + ex)
            batmlay = boftlay * math.exp(-tauprof[i + 1]) * (1.0 - ex)
            boftatm[i] = boftatm[i + 1] + batmlay
            tauprof[i] = tauprof[i + 1] + taulay[i + 1]
        if tauprof[0] < EXPMAX:
            boftbg = es * tk2b_mod(hvk, tk[0])
            bakgrnd = boftbg * math.exp(-tauprof[0])
            boftotl = bakgrnd + boftatm[0]
            boftmr = boftatm[0] / (1.0 - math.exp(-tauprof[0]))
        else:
            bakgrnd = 0.0
            boftotl = boftatm[0]
            boftmr = boftatm[0]
        return boftotl, boftatm, boftmr, tauprof, hvk, boft, bakgrnd
    boft[0] = tk2b_mod(hvk, tk[0])
    for i in range(1, nl):
        boft[i] = tk2b_mod(hvk, tk[i])
        ex = math.exp(-taulay[i])
        boftlay = (boft[i - 1] + boft[i] * ex) / (1.0 + ex)
        batmlay = boftlay * math.exp(-tauprof[i - 1]) * (1.0 - ex)
        boftatm[i] = boftatm[i - 1] + batmlay
        tauprof[i] = tauprof[i - 1] + taulay[i]
    if tauprof[nl - 1] < EXPMAX:
        boftbg = tk2b_mod(hvk, T_COSMIC)
        bakgrnd = boftbg * math.exp(-tauprof[nl - 1])
        boftotl = bakgrnd + boftatm[nl - 1]
        boftmr = boftatm[nl - 1] / (1.0 - math.exp(-tauprof[nl - 1]))
    else:
        bakgrnd = 0.0
        boftotl = boftatm[nl - 1]
        boftmr = boftatm[nl - 1]
    return boftotl, boftatm, boftmr, tauprof, hvk, boft, bakgrnd


def cloud_radiating_temperature(ibase, itop, hvk, tauprof, boftatm):
    """Mean radiating temperature of one cloud layer (column `tmrcld`)."""
    if tauprof[ibase] > EXPMAX:
        return 0.0
    batmcld = boftatm[itop] - boftatm[ibase]
    taucld = tauprof[itop] - tauprof[ibase]
    if taucld > EXPMAX:
        boftcld = batmcld * math.exp(tauprof[ibase])
    else:
        boftcld = batmcld * math.exp(tauprof[ibase]) / (1.0 - math.exp(-taucld))
    return bright(hvk, boftcld)


# --------------------------------------------------------------------------------------------
# pyrtlib.tb_spectrum.TbCloudRTE
# --------------------------------------------------------------------------------------------
COLUMNS = ("tbtotal", "tbatm", "tmr", "tmrcld", "tauwet", "taudry", "tauliq", "tauice")


class TbCloudRTE:
    """Restatement of `pyrtlib.tb_spectrum.TbCloudRTE` (call sites: PyRTlib_processing.py:123-127).

    z [km], p [mb], t [K], rh [fraction], frq [GHz], angles [deg elevation]; arrays ground-first.
    `execute()` returns a dict of 1-D arrays of length nf*nang, angle-major / frequency fastest
    (the row order of upstream's DataFrame) -- no pandas in the oracle.
    """

    def __init__(self, z, p, t, rh, frq, angles, o3n=None, amu=None, absmdl="", ray_tracing=False,
                 from_sat=True, cloudy=False):
        self.z = np.array(z, dtype=np.float64)          # copies: inputs are never mutated
        self.p = np.array(p, dtype=np.float64)
        self.tk = np.array(t, dtype=np.float64)
        self.rh = np.array(rh, dtype=np.float64)
        self.frq = np.atleast_1d(np.array(frq, dtype=np.float64))
        self.angles = np.atleast_1d(np.array(angles, dtype=np.float64))
        if o3n is not None or amu is not None:
            raise NotImplementedError("o3n / amu are outside the reference's path")
        self.ray_tracing = ray_tracing
        self.satellite = from_sat
        self.cloudy = cloudy
        self.emissivity = 1.0
        self.nl = len(self.z)
        self.nf = len(self.frq)
        self.nang = len(self.angles)
        self.z0 = self.z[0]
        self.z = self.z - self.z0
        self.denliq = np.zeros(self.nl)
        self.denice = np.zeros(self.nl)
        self.cldh = np.zeros((2, 0))
        self._mdl = None
        if absmdl:
            self.init_absmdl(absmdl)

    def init_absmdl(self, absmdl):
        self._mdl = model_tables.load(absmdl)

    def init_cloudy(self, cldh, denice, denliq):
        self.cloudy = True
        self.cldh = np.array(cldh, dtype=np.float64).reshape(2, -1) - self.z0
        self.denice = np.array(denice, dtype=np.float64)
        self.denliq = np.array(denliq, dtype=np.float64)

    def _cloud_levels(self):
        """Indices of the levels at cloud base/top (exact matches expected, as upstream)."""
        ibase, itop = [], []
        for l in range(self.cldh.shape[1]):
            ibase.append(int(np.argmin(np.abs(self.z - self.cldh[0, l]))))
            itop.append(int(np.argmin(np.abs(self.z - self.cldh[1, l]))))
        return ibase, itop

    def execute(self):
        if self._mdl is None:
            raise ValueError("init_absmdl() must be called before execute()")
        nl, nf, nang = self.nl, self.nf, self.nang
        out = {c: np.zeros((nf, nang)) for c in COLUMNS}
        self.ptaudry = np.zeros((nf, nang, nl))
        self.ptauwet = np.zeros((nf, nang, nl))
        self.ptauliq = np.zeros((nf, nang, nl))
        self.ptauice = np.zeros((nf, nang, nl))
        self.srho = np.zeros(nang)
        e, rho = vapor(self.tk, self.rh)
        for k in range(nang):
            if self.ray_tracing:
                _, _, refindx = refractivity(self.p, self.tk, e)
                ds = ray_tracing(self.z, refindx, self.angles[k], self.z0)
            else:
                amass = 1.0 / math.sin(self.angles[k] * math.pi / 180.0)
                ds = np.zeros(nl)
                for i in range(1, nl):
                    ds[i] = (self.z[i] - self.z[i - 1]) * amass
            self.srho[k], _ = exponential_integration(True, rho, ds, 0, nl, 0.1)
            for j in range(nf):
                awet, adry = clearsky_absorption(self._mdl, self.p, self.tk, e, self.frq[j])
                out["tauwet"][j, k], self.ptauwet[j, k, :] = exponential_integration(
                    True, awet, ds, 0, nl, 1)
                out["taudry"][j, k], self.ptaudry[j, k, :] = exponential_integration(
                    True, adry, ds, 0, nl, 1)
                if self.cloudy:
                    aliq, aice = cloudy_absorption(self._mdl, self.tk, self.denliq, self.denice,
                                                   self.frq[j])
                    out["tauliq"][j, k], self.ptauliq[j, k, :] = exponential_integration(
                        False, aliq, ds, 0, nl, 1)
                    out["tauice"][j, k], self.ptauice[j, k, :] = exponential_integration(
                        False, aice, ds, 0, nl, 1)
                ptaulay = (self.ptauwet[j, k, :] + self.ptaudry[j, k, :] + self.ptauice[j, k, :]
                           + self.ptauliq[j, k, :])
                boftotl, boftatm, boftmr, tauprof, hvk, _, _ = planck(
                    self.frq[j], self.tk, ptaulay, from_sat=self.satellite, es=self.emissivity)
                iend = 0 if self.satellite else nl - 1
                out["tbtotal"][j, k] = bright(hvk, boftotl)
                out["tbatm"][j, k] = bright(hvk, boftatm[iend])
                out["tmr"][j, k] = bright(hvk, boftmr)
                if self.cloudy and self.cldh.shape[1] > 0 and not self.satellite:
                    ibase, itop = self._cloud_levels()
                    out["tmrcld"][j, k] = cloud_radiating_temperature(ibase[0], itop[0], hvk,
                                                                      tauprof, boftatm)
        res = {c: out[c].T.reshape(nf * nang) for c in COLUMNS}
        res["angle"] = np.repeat(self.angles, nf)
        res["frq"] = np.tile(self.frq, nang)
        return res
