"""CPU restatement of the reference's cloud detection and adiabatic water content (SURVEY.md 8(f) N3,
"batched cloud detection").  TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows /root/reference/python_src/preproc/derive_cloud_water.py:
  clausius_clapeyron_liq/ice :49-64      calc_lwc :68-142      derive_cloud_features :146-363
(the relative-humidity layer detection of Nandan et al. 2022 with the reference's own list handling,
including its quirks: strict band limits, `z_array[j-1]` wrapping at j = 0, log of a non-positive
height above base).  Arrays are TOP-first like the reference's (`Level_*`, index 0 = top).
`cloud_bools` of the reference is bookkeeping that no output depends on; it is not restated.
Pinned bit-exact against outputs of the reference's own function (tests/golden/make_cloud_golden.py).
"""
import numpy as np

BANDS = ((92, 95, 84), (90, 93, 82), (88, 90, 78), (75, 80, 70))   # (min, max, inter) RH per band


def cc_liq(tc):
    return 610.78 * np.exp(2.5e6 / 462 * (1 / 273.15 - 1 / (273.15 + tc)))


def cc_ice(tc):
    return 610.78 * np.exp(2.840e6 / 462 * (1 / 273.15 - 1 / (273.15 + tc)))


def band(z):
    """Index of the altitude band of derive_cloud_features, or -1 on a band limit (strict tests)."""
    if z < 2000:
        return 0
    if 2000 < z < 6000:
        return 1
    if 6000 < z < 12000:
        return 2
    if 12000 < z:
        return 3
    return -1


def _idx(z_array, x):
    return int(np.nanargmin(np.abs(z_array - x)))


def calc_lwc(tops, bases, p, t, z):
    cp, L, R_L, gamma_d, gamma_s = 1003.5, 334944, 287.06, 9.76e-3, 6.5e-3
    n = len(t)
    lwc_m3, lwc_kg, iwc_m3, iwc_kg = (np.zeros(n) for _ in range(4))
    for base, top in zip(bases, tops):
        it, ib = _idx(z, top), _idx(z, base)
        if t[ib] > 273.15 and t[it] > 273.15:
            tgt = (lwc_m3, lwc_kg)
        elif t[ib] < 233.15 and t[it] < 233.15:
            tgt = (iwc_m3, iwc_kg)
        elif t[ib] > 233.15 and t[it] < 273.15:
            tgt = (lwc_m3, lwc_kg)
        else:
            continue
        for j in range(it, ib):
            rho = p[j] * 100 / R_L / t[j]
            dz = z[j - 1] - z[j]
            ad = rho * cp / L * (gamma_d - gamma_s) * dz
            dh = z[j] - base
            with np.errstate(all="ignore"):
                w = ad * (1.239 - 0.145 * np.log(dh))
            if w < 0:
                w = 0
            tgt[0][j] = w
            tgt[1][j] = w / rho
    lwp = np.abs(np.sum(lwc_m3 * np.gradient(z)))
    iwp = np.abs(np.sum(iwc_m3 * np.gradient(z)))
    return lwc_m3, lwc_kg, lwp, iwc_m3, iwc_kg, iwp


def derive_cloud_features(p, t, z, rh):
    p, t, z = (np.asarray(a, dtype=np.float64) for a in (p, t, z))
    rh = np.array(rh, dtype=np.float64, copy=True)
    for i, temp in enumerate(t):                      # 1) RH over ice below 0 C
        if temp < 273.15:
            rh[i] = rh[i] * cc_liq(temp - 273.15) / cc_ice(temp - 273.15)
    tops, bases = [], []
    in_cloud = False
    for i, zz in enumerate(z):                        # 2-4) preliminary layers above RH_min
        b = band(zz)
        if b < 0:
            continue
        if rh[i] > BANDS[b][0]:
            if not in_cloud:
                tops.append(zz)
            in_cloud = True
        else:
            if in_cloud:
                bases.append(zz)
            in_cloud = False
    if len(tops) == len(bases) + 1:
        bases.append(z[-1])
    # 5) bases below 500 m with a thickness under 400 m
    keep = [i for i, (b_, t_) in enumerate(zip(bases, tops)) if not (b_ < 500 and abs(b_ - t_) < 400)]
    bases, tops = [bases[i] for i in keep], [tops[i] for i in keep]
    # 6) RH_max must be reached inside the layer
    rm = []
    for i, (base, top) in enumerate(zip(bases, tops)):
        b = band(base)
        if b < 0:
            continue
        it, ib = _idx(z, top), _idx(z, base)
        if not np.any(BANDS[b][1] < rh[it:ib]):
            rm.append(i)
    for i in sorted(rm, reverse=True):
        bases.pop(i)
        tops.pop(i)
    # 7) connect layers
    rmb, rmt = [], []
    rh_inter = None
    for i, (base, top) in enumerate(zip(bases, tops)):
        b = band(base)
        if b >= 0:
            rh_inter = BANDS[b][2]
        if i != 0:
            ib, it = _idx(z, bases[i - 1]), _idx(z, top)
            if 1 == abs(ib - it):
                if rh[ib] > rh_inter:
                    rmb.append(i - 1)
                    rmt.append(i)
            elif abs(bases[i - 1] - top) < 300:
                rmb.append(i - 1)
                rmt.append(i)
            else:
                gap = rh[ib:it]
                with np.errstate(all="ignore"):
                    if gap.size > 0 and np.nanmin(gap) > rh_inter:
                        rmb.append(i - 1)
                        rmt.append(i)
    for i in sorted(rmb, reverse=True):
        bases.pop(i)
    for i in sorted(rmt, reverse=True):
        tops.pop(i)
    # 8) thinner than 100 m
    rm = [i for i, (base, top) in enumerate(zip(bases, tops)) if abs(base - top) < 100]
    for i in sorted(rm, reverse=True):
        bases.pop(i)
        tops.pop(i)
    out = calc_lwc(tops, bases, p, t, z)
    return out + (np.array(tops), np.array(bases))
