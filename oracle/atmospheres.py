"""AFGL US-standard atmosphere and the ppmv->RH helper chain (BASELINE.json config 1 input).
TEST INFRASTRUCTURE (see oracle/__init__.py).

Restates `pyrtlib.climatology.AtmosphericProfiles.gl_atm(US_STANDARD)` and
`pyrtlib.utils.ppmv2gkg / mr2rh` -- the chain the reference uses at
`/root/reference/python_src/preproc/preprocessing4all.py:503-505` (there with MIDLATITUDE_SUMMER).
Values are the AFGL-TR-86-0110 tables as recorded in SURVEY.md Appendix C (from memory; a wrong digit
changes the input profile, not GPU-vs-oracle parity).
"""
import numpy as np

Z_KM = np.array(list(range(0, 26)) + [27.5, 30.0, 32.5, 35.0, 37.5, 40.0, 42.5, 45.0, 47.5, 50.0]
                + [55.0, 60.0, 65.0, 70.0, 75.0, 80.0, 85.0, 90.0, 95.0, 100.0, 105.0, 110.0, 115.0,
                   120.0], dtype=np.float64)

P_MB = np.array([1013., 898.8, 795.0, 701.2, 616.6, 540.5, 472.2, 411.1, 356.5, 308.0, 265.0, 227.0,
                 194.0, 165.8, 141.7, 121.1, 103.5, 88.50, 75.65, 64.67, 55.29, 47.29, 40.47, 34.67,
                 29.72, 25.49, 17.43, 11.97, 8.010, 5.746, 4.150, 2.871, 2.060, 1.491, 1.090, 0.7978,
                 0.4250, 0.2190, 0.1090, 0.05220, 0.02400, 0.01050, 4.460e-3, 1.840e-3, 7.600e-4,
                 3.200e-4, 1.450e-4, 7.100e-5, 4.010e-5, 2.540e-5])

T_K = np.array([288.2, 281.7, 275.2, 268.7, 262.2, 255.7, 249.2, 242.7, 236.2, 229.7, 223.3, 216.8,
                216.7, 216.7, 216.7, 216.7, 216.7, 216.7, 216.7, 216.7, 216.7, 217.6, 218.6, 219.6,
                220.6, 221.6, 224.0, 226.5, 230.0, 236.5, 242.9, 250.4, 257.3, 264.2, 270.6, 270.7,
                260.8, 247.0, 233.3, 219.6, 208.4, 198.6, 188.9, 186.9, 188.4, 195.1, 208.8, 240.0,
                300.0, 360.0])

H2O_PPMV = np.array([7745., 6071., 4631., 3182., 2158., 1397., 925.4, 572.0, 366.7, 158.3, 69.96,
                     36.13, 19.06, 10.85, 5.927, 5.000, 3.950, 3.850, 3.825, 3.850, 3.900, 3.975,
                     4.065, 4.200, 4.300, 4.425, 4.575, 4.725, 4.825, 4.900, 4.950, 5.025, 5.150,
                     5.225, 5.250, 5.225, 5.100, 4.750, 4.200, 3.500, 2.825, 2.050, 1.330, 0.850,
                     0.540, 0.400, 0.340, 0.280, 0.240, 0.200])

assert len(Z_KM) == len(P_MB) == len(T_K) == len(H2O_PPMV) == 50


def ppmv2gkg(ppmv, molar_mass=18.01528):
    return ppmv * (molar_mass / 28.9644) * 1e-3


def goff_gratch_water(tk):
    y = 373.16 / tk
    es = (-7.90298 * (y - 1.0) + 5.02808 * np.log10(y)
          - 1.3816e-07 * (10 ** (11.344 * (1.0 - (1.0 / y))) - 1.0)
          + 0.0081328 * (10 ** (-3.49149 * (y - 1.0)) - 1.0) + np.log10(1013.246))
    return 10.0 ** es


def mr2rh(p, tk, w_gkg):
    """(rh from mixing ratio, rh from vapour pressure), both as fractions x100 = % like upstream."""
    es = goff_gratch_water(tk)
    wsat = 622.0 * es / p  # upstream satmix(): es/p form (keeps RH positive above 60 km)
    e = p * w_gkg / (622.0 + w_gkg)
    return 100.0 * w_gkg / wsat, 100.0 * e / es


def us_standard(nlev=None):
    """z[km], p[mb], t[K], rh[fraction], ground-first -- what config 1 feeds TbCloudRTE."""
    gkg = ppmv2gkg(H2O_PPMV)
    rh = mr2rh(P_MB, T_K, gkg)[0] / 100.0
    z, p, t = Z_KM.copy(), P_MB.copy(), T_K.copy()
    if nlev is not None:
        z, p, t, rh = z[:nlev], p[:nlev], t[:nlev], rh[:nlev]
    return z, p, t, rh
