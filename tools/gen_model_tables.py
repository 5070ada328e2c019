#!/usr/bin/env python3
"""Generate the absorption-model descriptor files.

Outputs (all committed; re-run this tool after editing a table):
  mwr_fast_forward_operators_and_lbls_b200/data/models/<MODEL>.json   one descriptor per model string
  mwr_fast_forward_operators_and_lbls_b200/csrc/mwrt_builtin_models.inc  the same tables as C initialisers

PROVENANCE -- READ THIS.
The reference (/root/reference) calls the third-party package pyrtlib==1.1.1
(requirements.txt:278; PyRTlib_processing.py:26-29) whose source and spectroscopic line lists are NOT
in the reference tree and cannot be installed offline.  Every number below is a
RECOLLECTION of P. W. Rosenkranz's public-domain Fortran line lists (abh2o/h2o_list/h2o_sdlist,
o2abs/o2_list) as re-packaged by PyRTlib, written down without access to either source.  Leading
digits are believed right; trailing digits, and the second-order O2 mixing coefficients in particular,
are UNVERIFIED.  GPU-vs-oracle parity does not depend on these digits (both sides read the same
descriptor); agreement with a real PyRTlib run does.  `baseline/probe_pyrtlib.py` dumps the true
tables from an installed pyrtlib into this same JSON schema so they can be swapped in.

Model strings are the ones the reference uses (PyRTlib_processing.py:90,122,130,138,146) plus the
R21SD/R22SD/R23/R23SD members of the upstream family (BASELINE.json config 1 names R22SD).
"""
import copy
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(HERE, "..", "mwr_fast_forward_operators_and_lbls_b200")
OUT_JSON = os.path.join(PKG, "data", "models")
OUT_INC = os.path.join(PKG, "csrc", "mwrt_builtin_models.inc")

# ----------------------------------------------------------------------------------------------
# H2O line lists
# ----------------------------------------------------------------------------------------------
# 1998-era 15-line list.  Widths in GHz/mb.
H2O_98 = dict(
    fl=[22.2351, 183.3101, 321.2256, 325.1529, 380.1974, 439.1508, 443.0183, 448.0011, 470.8890,
        474.6891, 488.4911, 556.9360, 620.7008, 752.0332, 916.1712],
    s1=[0.1310e-13, 0.2273e-11, 0.8036e-13, 0.2694e-11, 0.2438e-10, 0.2179e-11, 0.4624e-12,
        0.2562e-10, 0.8369e-12, 0.3263e-11, 0.6659e-12, 0.1531e-08, 0.1707e-10, 0.1011e-08,
        0.4227e-10],
    b2=[2.144, 0.668, 6.179, 1.541, 1.048, 3.595, 5.048, 1.405, 3.597, 2.379, 2.852, 0.159, 2.391,
        0.396, 1.441],
    w0=[0.00281, 0.00281, 0.00230, 0.00278, 0.00287, 0.00210, 0.00186, 0.00263, 0.00215, 0.00236,
        0.00260, 0.00321, 0.00244, 0.00306, 0.00267],
    x=[0.69, 0.64, 0.67, 0.68, 0.54, 0.63, 0.60, 0.66, 0.66, 0.65, 0.69, 0.69, 0.71, 0.68, 0.70],
    w0s=[0.01349, 0.01491, 0.01080, 0.01350, 0.01541, 0.00900, 0.00788, 0.01275, 0.00983, 0.01095,
         0.01313, 0.01320, 0.01140, 0.01253, 0.01275],
    xs=[0.61, 0.85, 0.54, 0.74, 0.89, 0.52, 0.50, 0.67, 0.65, 0.64, 0.72, 1.00, 0.68, 0.84, 0.78],
)

# 2003 (Tretyakov et al. widths for the 22 and 183 GHz lines)
H2O_03 = copy.deepcopy(H2O_98)
H2O_03["w0"][0] = 0.002656
H2O_03["w0"][1] = 0.00281
H2O_03["w0s"][0] = 0.0127488
H2O_03["w0s"][1] = 0.01491

# 2016/2017 15-line list: updated 22 GHz intensity and widths, shift given as a ratio to the width
H2O_17 = copy.deepcopy(H2O_03)
H2O_17["s1"][0] = 0.1321e-13
H2O_17["x"] = [0.76, 0.77, 0.67, 0.68, 0.54, 0.63, 0.60, 0.66, 0.66, 0.65, 0.69, 0.69, 0.71, 0.68,
               0.70]
H2O_17["xs"] = [0.62, 0.85, 0.54, 0.74, 0.89, 0.52, 0.50, 0.67, 0.65, 0.64, 0.72, 1.00, 0.68, 0.84,
                0.78]
H2O_17["sr"] = [0.0, -0.017] + [0.0] * 13

# 2019+ 16-line list (h2o_sdlist), reference temperature 296 K; widths/shifts stored in GHz/mb
_k = 1.0e-3
H2O_19 = dict(
    fl=[22.235080, 183.310087, 321.225630, 325.152888, 380.197353, 439.150807, 443.018343,
        448.001085, 470.888999, 474.689092, 488.490108, 556.935985, 620.700807, 658.006072,
        752.033113, 916.171582],
    s1=[0.1335e-13, 0.2319e-11, 0.7657e-13, 0.2721e-11, 0.2477e-10, 0.2137e-11, 0.4440e-12,
        0.2588e-10, 0.8196e-12, 0.3268e-11, 0.6628e-12, 0.1570e-08, 0.1700e-10, 0.9033e-12,
        0.1035e-08, 0.4275e-10],
    b2=[2.172, 0.677, 6.262, 1.561, 1.062, 3.643, 5.116, 1.424, 3.645, 2.411, 2.890, 0.161, 2.423,
        7.921, 0.402, 1.461],
    w0=[_k * v for v in [2.699, 2.945, 2.426, 2.847, 2.868, 2.055, 1.819, 2.612, 2.169, 2.366,
                         2.616, 3.115, 2.468, 3.154, 3.114, 2.695]],
    x=[0.76, 0.77, 0.73, 0.64, 0.54, 0.69, 0.70, 0.70, 0.73, 0.71, 0.75, 0.75, 0.79, 0.73, 0.77,
       0.79],
    w0s=[_k * v for v in [13.29, 14.78, 10.65, 13.95, 14.40, 9.06, 7.96, 13.01, 9.70, 11.24, 13.58,
                          14.24, 11.94, 13.84, 13.58, 13.55]],
    xs=[0.62, 0.77, 0.54, 0.74, 0.89, 0.52, 0.50, 0.67, 0.65, 0.64, 0.72, 1.00, 0.75, 1.00, 0.84,
        0.48],
    sh=[_k * v for v in [-0.033, -0.072, -0.143, -0.013, -0.074, 0.051, 0.140, -0.116, 0.061,
                         -0.027, -0.065, 0.187, 0.000, 0.176, 0.162, 0.000]],
    xh=[2.6, 1.8] + [0.0] * 14,
    shs=[_k * v for v in [0.814, 0.173, 0.278, 1.325, 0.240, 0.165, -0.229, -0.615, -0.465, -0.720,
                          -0.360, -1.693, 0.687, -1.496, -0.878, 0.521]],
    xhs=[0.0] * 12 + [0.92, 0.0, 0.0, 0.47],
    w2=[_k * 0.435, _k * 0.407] + [0.0] * 14,
    w2s=[_k * 1.91, _k * 1.46] + [0.0] * 14,
)

# 2021/2022 speed-dependent list: Koshelev et al. (2021) shift temperature dependence for the two
# diagnostic lines (aair/aself) and speed dependence of the shift (d2).
H2O_22 = copy.deepcopy(H2O_19)
H2O_22["s1"][0] = 0.1335e-13
H2O_22["sh"][1] = _k * -0.096
H2O_22["aair"] = [0.0, 0.0] + [0.0] * 14
H2O_22["aself"] = [12.6, 0.0] + [0.0] * 14
H2O_22["xhs"] = [3.6, 0.0] + [0.0] * 10 + [0.92, 0.0, 0.0, 0.47]
H2O_22["shs"][0] = _k * 1.099
H2O_22["d2"] = [0.0, _k * -0.016] + [0.0] * 14
H2O_22["d2s"] = [0.0, _k * 0.16] + [0.0] * 14

H2O_24 = copy.deepcopy(H2O_22)
H2O_24["w0"][0] = _k * 2.704
H2O_24["w0"][1] = _k * 2.952


def h2o_block(lines, *, reftline, cf, xcf, cs, xcs, pvap_div, den_const, line_const, sd):
    n = len(lines["fl"])
    blk = dict(reftline=reftline, reftcon=300.0, cf=cf, xcf=xcf, cs=cs, xcs=xcs, pvap_div=pvap_div,
               den_const=den_const, line_const=line_const, sd=bool(sd))
    for key in ("fl", "s1", "b2", "w0", "x", "w0s", "xs", "sr", "sh", "xh", "shs", "xhs", "aair",
                "aself", "w2", "w2s", "d2", "d2s"):
        v = list(lines.get(key, [0.0] * n))
        assert len(v) == n, (key, len(v), n)
        blk[key] = [float(t) for t in v]
    if not sd:
        blk["w2"] = [0.0] * n
        blk["w2s"] = [0.0] * n
        blk["d2"] = [0.0] * n
        blk["d2s"] = [0.0] * n
    return blk


# ----------------------------------------------------------------------------------------------
# O2 line lists
# ----------------------------------------------------------------------------------------------
O2_98 = dict(
    f=[118.7503, 56.2648, 62.4863, 58.4466, 60.3061, 59.5910, 59.1642, 60.4348, 58.3239, 61.1506,
       57.6125, 61.8002, 56.9682, 62.4112, 56.3634, 62.9980, 55.7838, 63.5685, 55.2214, 64.1278,
       54.6712, 64.6789, 54.1300, 65.2241, 53.5957, 65.7648, 53.0669, 66.3021, 52.5424, 66.8368,
       52.0214, 67.3696, 51.5034, 67.9009, 368.4984, 424.7632, 487.2494, 715.3931, 773.8397,
       834.1458],
    s300=[0.2936e-14, 0.8079e-15, 0.2480e-14, 0.2228e-14, 0.3351e-14, 0.3292e-14, 0.3721e-14,
          0.3891e-14, 0.3640e-14, 0.4005e-14, 0.3227e-14, 0.3715e-14, 0.2627e-14, 0.3156e-14,
          0.1982e-14, 0.2477e-14, 0.1391e-14, 0.1808e-14, 0.9124e-15, 0.1230e-14, 0.5603e-15,
          0.7842e-15, 0.3228e-15, 0.4689e-15, 0.1748e-15, 0.2632e-15, 0.8898e-16, 0.1389e-15,
          0.4264e-16, 0.6899e-16, 0.1924e-16, 0.3229e-16, 0.8191e-17, 0.1423e-16, 0.6494e-15,
          0.7083e-14, 0.3025e-14, 0.1835e-14, 0.1158e-13, 0.3993e-14],
    be=[0.009, 0.015, 0.083, 0.084, 0.212, 0.212, 0.391, 0.391, 0.626, 0.626, 0.915, 0.915, 1.260,
        1.260, 1.660, 1.665, 2.119, 2.115, 2.624, 2.625, 3.194, 3.194, 3.814, 3.814, 4.484, 4.484,
        5.224, 5.224, 6.004, 6.004, 6.844, 6.844, 7.744, 7.744, 0.048, 0.044, 0.049, 0.145, 0.141,
        0.145],
    w300=[1.63, 1.646, 1.468, 1.449, 1.382, 1.360, 1.319, 1.297, 1.266, 1.248, 1.221, 1.207, 1.181,
          1.171, 1.144, 1.139, 1.110, 1.108, 1.079, 1.078, 1.05, 1.05, 1.02, 1.02, 1.00, 1.00, 0.97,
          0.97, 0.94, 0.94, 0.92, 0.92, 0.89, 0.89, 1.92, 1.92, 1.92, 1.81, 1.81, 1.81],
    y0=[-0.0233, 0.2408, -0.3486, 0.5227, -0.5430, 0.5877, -0.3970, 0.3237, -0.1348, 0.0311, 0.0725,
        -0.1663, 0.2832, -0.3629, 0.3970, -0.4599, 0.4695, -0.5199, 0.5187, -0.5597, 0.5903,
        -0.6246, 0.6656, -0.6942, 0.7086, -0.7325, 0.7348, -0.7546, 0.7702, -0.7864, 0.8083,
        -0.8210, 0.8439, -0.8529, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0],
    y1=[0.0079, -0.0978, 0.0844, -0.1273, 0.0699, -0.0776, 0.2309, -0.2825, 0.0436, -0.0584, 0.6056,
        -0.6619, 0.6451, -0.6759, 0.6547, -0.6675, 0.6135, -0.6139, 0.2952, -0.2895, 0.2654,
        -0.2590, 0.3750, -0.3680, 0.5085, -0.5002, 0.6206, -0.6091, 0.6526, -0.6393, 0.6640,
        -0.6475, 0.6729, -0.6545, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0],
)

_F49 = [118.7503, 56.2648, 62.4863, 58.4466, 60.3061, 59.5910, 59.1642, 60.4348, 58.3239, 61.1506,
        57.6125, 61.8002, 56.9682, 62.4112, 56.3634, 62.9980, 55.7838, 63.5685, 55.2214, 64.1278,
        54.6712, 64.6789, 54.1300, 65.2241, 53.5958, 65.7648, 53.0669, 66.3021, 52.5424, 66.8368,
        52.0214, 67.3696, 51.5034, 67.9009, 50.9877, 68.4310, 50.4742, 68.9603, 233.9461, 368.4982,
        401.7398, 424.7630, 487.2493, 566.8956, 715.3929, 731.1866, 773.8395, 834.1455, 895.0710]
_S49 = [0.2906e-14, 0.7957e-15, 0.2444e-14, 0.2194e-14, 0.3301e-14, 0.3243e-14, 0.3664e-14,
        0.3834e-14, 0.3588e-14, 0.3947e-14, 0.3179e-14, 0.3661e-14, 0.2590e-14, 0.3111e-14,
        0.1954e-14, 0.2443e-14, 0.1373e-14, 0.1784e-14, 0.9013e-15, 0.1217e-14, 0.5545e-15,
        0.7766e-15, 0.3201e-15, 0.4651e-15, 0.1738e-15, 0.2619e-15, 0.8880e-16, 0.1387e-15,
        0.4272e-16, 0.6923e-16, 0.1939e-16, 0.3255e-16, 0.8301e-17, 0.1445e-16, 0.3356e-17,
        0.6049e-17, 0.1280e-17, 0.2394e-17, 0.3287e-16, 0.6463e-15, 0.1334e-16, 0.7049e-14,
        0.3011e-14, 0.1797e-16, 0.1826e-14, 0.2193e-16, 0.1153e-13, 0.3974e-14, 0.2512e-16]
_BE49 = [0.010, 0.014, 0.083, 0.083, 0.207, 0.207, 0.387, 0.387, 0.621, 0.621, 0.910, 0.910, 1.255,
         1.255, 1.654, 1.654, 2.109, 2.108, 2.618, 2.617, 3.182, 3.181, 3.800, 3.800, 4.474, 4.473,
         5.201, 5.200, 5.983, 5.982, 6.819, 6.818, 7.709, 7.708, 8.653, 8.652, 9.651, 9.650, 0.019,
         0.048, 0.045, 0.044, 0.049, 0.084, 0.145, 0.136, 0.141, 0.145, 0.201]
_W49 = [1.688, 1.703, 1.513, 1.491, 1.415, 1.408, 1.353, 1.339, 1.295, 1.292, 1.262, 1.263, 1.223,
        1.217, 1.189, 1.174, 1.134, 1.134, 1.089, 1.088, 1.037, 1.038, 0.996, 0.996, 0.955, 0.955,
        0.906, 0.906, 0.858, 0.858, 0.811, 0.811, 0.764, 0.764, 0.717, 0.717, 0.669, 0.669, 2.78,
        1.64, 1.64, 1.64, 1.60, 1.60, 1.60, 1.60, 1.62, 1.47, 1.47]
_Z11 = [0.0] * 11

# Tretyakov et al. (2005) first-order mixing on the 49-line list (R16/R17)
O2_17 = dict(
    f=_F49, s300=_S49, be=_BE49, w300=_W49,
    y0=[-0.0360, 0.2547, -0.3655, 0.5495, -0.5696, 0.6181, -0.4252, 0.3517, -0.1496, 0.0430, 0.0640,
        -0.1605, 0.2906, -0.3730, 0.4169, -0.4819, 0.4963, -0.5481, 0.5512, -0.5931, 0.6212,
        -0.6558, 0.6920, -0.7208, 0.7312, -0.7550, 0.7555, -0.7751, 0.7914, -0.8073, 0.8307,
        -0.8431, 0.8676, -0.8761, 0.9046, -0.9092, 0.9416, -0.9423] + _Z11,
    y1=[0.0079, -0.0978, 0.0844, -0.1273, 0.0699, -0.0776, 0.2309, -0.2825, 0.0436, -0.0584, 0.6056,
        -0.6619, 0.6451, -0.6759, 0.6547, -0.6675, 0.6135, -0.6139, 0.2952, -0.2895, 0.2654,
        -0.2590, 0.3750, -0.3680, 0.5085, -0.5002, 0.6206, -0.6091, 0.6526, -0.6393, 0.6640,
        -0.6475, 0.6729, -0.6545, 0.680, -0.660, 0.685, -0.665] + _Z11,
)

# Makarov et al. (2011) second-order mixing (R19/R19SD).  Trailing digits UNVERIFIED.
O2_19 = dict(
    f=_F49, s300=_S49, be=_BE49, w300=_W49,
    y0=[-0.041, 0.277, -0.372, 0.559, -0.573, 0.618, -0.366, 0.278, -0.089, -0.021, 0.060, -0.152,
        0.216, -0.293, 0.373, -0.436, 0.491, -0.542, 0.571, -0.613, 0.636, -0.670, 0.690, -0.718,
        0.740, -0.763, 0.788, -0.807, 0.834, -0.849, 0.876, -0.887, 0.915, -0.922, 0.950, -0.955,
        0.987, -0.988] + _Z11,
    y1=[0.000, 0.110, -0.009, 0.007, 0.049, -0.100, 0.260, -0.346, 0.364, -0.422, 0.315, -0.341,
        0.483, -0.503, 0.598, -0.610, 0.630, -0.633, 0.613, -0.611, 0.570, -0.564, 0.580, -0.570,
        0.610, -0.600, 0.640, -0.620, 0.650, -0.640, 0.660, -0.640, 0.660, -0.640, 0.660, -0.640,
        0.650, -0.620] + _Z11,
    g0=[-0.000695, -0.090, -0.103, -0.239, -0.172, -0.171, 0.028, 0.150, 0.132, 0.170, 0.087, 0.069,
        0.083, 0.067, 0.007, 0.016, -0.021, -0.066, -0.095, -0.115, -0.118, -0.140, -0.173, -0.186,
        -0.217, -0.227, -0.234, -0.242, -0.266, -0.272, -0.301, -0.304, -0.334, -0.333, -0.361,
        -0.358, -0.348, -0.344] + _Z11,
    g1=[0.000, -0.042, 0.004, 0.025, 0.083, 0.167, 0.178, 0.223, 0.054, 0.003, 0.002, -0.044, -0.019,
        -0.054, -0.177, -0.208, -0.294, -0.334, -0.368, -0.386, -0.374, -0.384, -0.387, -0.389,
        -0.423, -0.422, -0.460, -0.460, -0.510, -0.500, -0.550, -0.540, -0.580, -0.560, -0.620,
        -0.590, -0.680, -0.650] + _Z11,
    dnu0=[-0.00028, 0.00596, -0.01950, 0.03200, -0.04750, 0.05410, -0.02320, 0.01550, 0.00070,
          -0.00860, -0.01120, 0.01320, -0.00210, 0.00400, 0.00510, -0.00290, 0.00640, -0.00430,
          0.00660, -0.00470, 0.00610, -0.00440, 0.00520, -0.00360, 0.00420, -0.00270, 0.00330,
          -0.00190, 0.00250, -0.00120, 0.00170, -0.00050, 0.00100, 0.00010, 0.00040, 0.00070,
          -0.00020, 0.00130] + _Z11,
    dnu1=[-0.00039, 0.00850, -0.00550, 0.00120, 0.01440, -0.02150, 0.02140, -0.02560, 0.01380,
          -0.01290, 0.00240, -0.00160, -0.00230, 0.00330, -0.00580, 0.00690, -0.00790, 0.00900,
          -0.00860, 0.00970, -0.00850, 0.00960, -0.00790, 0.00900, -0.00710, 0.00820, -0.00620,
          0.00730, -0.00530, 0.00640, -0.00440, 0.00550, -0.00350, 0.00460, -0.00260, 0.00370,
          -0.00170, 0.00280] + _Z11,
)

# Makarov et al. (2020) update (R20 ...): revised widths of the fine-structure lines and mixing.
O2_20 = copy.deepcopy(O2_19)
O2_20["w300"] = [1.685, 1.703, 1.513, 1.495, 1.433, 1.408, 1.353, 1.353, 1.303, 1.319, 1.262, 1.265,
                 1.238, 1.217, 1.207, 1.207, 1.137, 1.137, 1.101, 1.101, 1.037, 1.038, 0.996, 0.996,
                 0.955, 0.955, 0.906, 0.906, 0.858, 0.858, 0.811, 0.811, 0.764, 0.764, 0.717, 0.717,
                 0.669, 0.669, 2.78, 1.64, 1.64, 1.64, 1.60, 1.60, 1.60, 1.60, 1.62, 1.47, 1.47]
O2_20["y0"] = [-0.041, 0.277, -0.373, 0.560, -0.573, 0.618, -0.366, 0.278, -0.089, -0.021, 0.0599,
               -0.152, 0.216, -0.293, 0.374, -0.436, 0.491, -0.542, 0.571, -0.613, 0.636, -0.670,
               0.690, -0.718, 0.740, -0.763, 0.788, -0.807, 0.834, -0.849, 0.876, -0.887, 0.915,
               -0.922, 0.950, -0.955, 0.987, -0.988] + _Z11


def o2_block(lines, *, x, pvap_div, h2o_broad, nonres, abs_num, abs_den, iso, mixing, dens118,
             sd118):
    n = len(lines["f"])
    blk = dict(x=x, wb300=0.56, pvap_div=pvap_div, h2o_broad=h2o_broad, nonres=nonres,
               abs_num=abs_num, abs_den=abs_den, iso=iso, mixing=int(mixing),
               dens118=bool(dens118), sd118=bool(sd118))
    for key in ("f", "s300", "be", "w300", "y0", "y1", "g0", "g1", "dnu0", "dnu1"):
        v = list(lines.get(key, [0.0] * n))
        assert len(v) == n, (key, len(v), n)
        blk[key] = [float(t) for t in v]
    if mixing == 1:
        for key in ("g0", "g1", "dnu0", "dnu1"):
            blk[key] = [0.0] * n
    return blk


OLD_O2 = dict(x=0.8, pvap_div=217.0, h2o_broad=1.1, nonres=1.6e-17, abs_num=0.5034e12,
              abs_den=3.14159, iso=1.0, mixing=1, sd118=False)
NEW_O2 = dict(x=0.754, pvap_div=216.68, h2o_broad=1.2, nonres=1.584e-17, abs_num=1.6097e11,
              abs_den=1.0, mixing=2, dens118=False)

N2_98 = dict(l=6.4e-14, m=3.55, n=1.0, fdep=False)
N2_03 = dict(l=6.5e-14, m=3.6, n=1.29, fdep=True)
N2_16 = dict(l=6.5e-14, m=3.6, n=1.34, fdep=True)
N2_20 = dict(l=9.95e-14, m=3.22, n=1.0, fdep=True)

UNVERIFIED = ("Recollection of Rosenkranz line lists as packaged by PyRTlib 1.1.x; written without "
              "access to either source (SURVEY.md 8(c)); trailing digits UNVERIFIED - parity with "
              "real PyRTlib is unpinned.")

MODELS = {}


def add(name, note, h2o, o2, n2, liq):
    MODELS[name] = dict(name=name, provenance=UNVERIFIED, note=note, h2o=h2o, o2=o2, n2=n2,
                        liq=dict(model=liq))


add("R98", "Rosenkranz 1998: 15 H2O lines, 40 O2 lines (Liebe 1992, first-order mixing), Liebe-91 liquid",
    h2o_block(H2O_98, reftline=300.0, cf=5.43e-10, xcf=3.0, cs=1.8e-8, xcs=7.5, pvap_div=217.0,
              den_const=3.335e16, line_const=0.3183e-4, sd=False),
    o2_block(O2_98, dens118=True, **OLD_O2), N2_98, "liebe91")
add("R03", "R98 with Tretyakov et al. (2003) 22/183 GHz widths",
    h2o_block(H2O_03, reftline=300.0, cf=5.43e-10, xcf=3.0, cs=1.8e-8, xcs=7.5, pvap_div=217.0,
              den_const=3.335e16, line_const=0.3183e-4, sd=False),
    o2_block(O2_98, dens118=True, **OLD_O2), N2_03, "liebe91")
add("R16", "Rosenkranz 2016: Turner et al. (2009) continuum, Tretyakov et al. (2005) O2",
    h2o_block(H2O_17, reftline=300.0, cf=5.96e-10, xcf=3.0, cs=1.42e-8, xcs=7.5, pvap_div=217.0,
              den_const=3.335e16, line_const=0.3183e-4, sd=False),
    o2_block(O2_17, dens118=False, **OLD_O2), N2_16, "liebe91")
add("R17", "Rosenkranz 2017: as R16 with revised constants and Rosenkranz (2015) liquid permittivity",
    h2o_block(H2O_17, reftline=300.0, cf=5.96e-10, xcf=3.0, cs=1.42e-8, xcs=7.5, pvap_div=217.0,
              den_const=3.344e16, line_const=0.3183e-4, sd=False),
    o2_block(O2_17, dens118=False, **OLD_O2), N2_16, "rosenkranz15")
_h19 = dict(reftline=296.0, cf=5.919e-10, xcf=3.0, cs=1.416e-8, xcs=7.5, pvap_div=216.68,
            den_const=3.344e16, line_const=3.1831e-5)
add("R19", "Rosenkranz 2019: 16 H2O lines (296 K), Makarov et al. (2011) second-order O2 mixing",
    h2o_block(H2O_19, sd=False, **_h19),
    o2_block(O2_19, iso=1.0, sd118=False, **NEW_O2), N2_16, "rosenkranz15")
add("R19SD", "R19 with speed-dependent 22/183 GHz H2O line shape",
    h2o_block(H2O_19, sd=True, **_h19),
    o2_block(O2_19, iso=1.0, sd118=False, **NEW_O2), N2_16, "rosenkranz15")
add("R20", "Rosenkranz 2020: Makarov et al. (2020) O2, 1.004 isotopologue factor, new N2",
    h2o_block(H2O_19, sd=False, **_h19),
    o2_block(O2_20, iso=1.004, sd118=False, **NEW_O2), N2_20, "rosenkranz15")
add("R20SD", "R20 with speed-dependent 22/183 GHz H2O line shape",
    h2o_block(H2O_19, sd=True, **_h19),
    o2_block(O2_20, iso=1.004, sd118=True, **NEW_O2), N2_20, "rosenkranz15")
_h22 = dict(_h19, cf=5.9197e-10, cs=1.4162e-8)
add("R21SD", "Rosenkranz 2021 SD: Koshelev et al. (2021) shift temperature dependence",
    h2o_block(H2O_22, sd=True, **_h22),
    o2_block(O2_20, iso=1.004, sd118=True, **NEW_O2), N2_20, "rosenkranz15")
add("R22SD", "Rosenkranz 2022 SD (BASELINE.json config 1)",
    h2o_block(H2O_22, sd=True, **_h22),
    o2_block(O2_20, iso=1.004, sd118=True, **NEW_O2), N2_20, "rosenkranz15")
add("R23", "Rosenkranz 2023 (non-SD variant of the 2022 list)",
    h2o_block(H2O_22, sd=False, **_h22),
    o2_block(O2_20, iso=1.004, sd118=False, **NEW_O2), N2_20, "rosenkranz15")
add("R23SD", "Rosenkranz 2023 SD",
    h2o_block(H2O_22, sd=True, **_h22),
    o2_block(O2_20, iso=1.004, sd118=True, **NEW_O2), N2_20, "rosenkranz15")
add("R24", "Rosenkranz 2024 - the reference's analysis baseline (PyRTlib_processing.py:128-135)",
    h2o_block(H2O_24, sd=True, **_h22),
    o2_block(O2_20, iso=1.004, sd118=True, **NEW_O2), N2_20, "rosenkranz15")

# ----------------------------------------------------------------------------------------------
# emit
# ----------------------------------------------------------------------------------------------
H2O_COLS = ["fl", "s1", "b2", "w0", "x", "w0s", "xs", "sr", "sh", "xh", "shs", "xhs", "aair", "aself",
            "w2", "w2s", "d2", "d2s"]
O2_COLS = ["f", "s300", "be", "w300", "y0", "y1", "g0", "g1", "dnu0", "dnu1"]


def c_double(v):
    return repr(float(v))


def emit_inc(path):
    out = ["// GENERATED by tools/gen_model_tables.py -- do not edit.  See that file for provenance:",
           "// every number is an UNVERIFIED recollection of the Rosenkranz/PyRTlib tables.",
           "// Layout matches struct mwrt_model_desc in include/mwrt.h; line tables are column-major",
           "// (one array of n_lines doubles per column, columns in MWRT_H2O_* / MWRT_O2_* order).", ""]
    names = list(MODELS)
    for name in names:
        m = MODELS[name]
        h, o = m["h2o"], m["o2"]
        out.append(f"static const double k_{name}_h2o[] = {{")
        for col in H2O_COLS:
            out.append("  " + ", ".join(c_double(v) for v in h[col]) + ",  // " + col)
        out.append("};")
        out.append(f"static const double k_{name}_o2[] = {{")
        for col in O2_COLS:
            out.append("  " + ", ".join(c_double(v) for v in o[col]) + ",  // " + col)
        out.append("};")
    out.append("")
    out.append("static const mwrt_builtin_model k_builtin_models[] = {")
    for name in names:
        m = MODELS[name]
        h, o, n2 = m["h2o"], m["o2"], m["n2"]
        liq = 0 if m["liq"]["model"] == "liebe91" else 1
        out.append(
            f'  {{"{name}", {{ {len(h["fl"])}, {len(o["f"])}, {int(h["sd"])}, {o["mixing"]}, '
            f'{int(o["dens118"])}, {int(o["sd118"])}, {int(n2["fdep"])}, {liq},\n'
            f'    {c_double(h["reftline"])}, {c_double(h["reftcon"])}, {c_double(h["cf"])}, '
            f'{c_double(h["xcf"])}, {c_double(h["cs"])}, {c_double(h["xcs"])}, '
            f'{c_double(h["pvap_div"])}, {c_double(h["den_const"])}, {c_double(h["line_const"])},\n'
            f'    {c_double(o["x"])}, {c_double(o["wb300"])}, {c_double(o["pvap_div"])}, '
            f'{c_double(o["h2o_broad"])}, {c_double(o["nonres"])}, {c_double(o["abs_num"])}, '
            f'{c_double(o["abs_den"])}, {c_double(o["iso"])},\n'
            f'    {c_double(n2["l"])}, {c_double(n2["m"])}, {c_double(n2["n"])},\n'
            f'    k_{name}_h2o, k_{name}_o2 }} }},')
    out.append("};")
    out.append(f"static const int k_n_builtin_models = {len(names)};")
    with open(path, "w") as fh:
        fh.write("\n".join(out) + "\n")


def main():
    os.makedirs(OUT_JSON, exist_ok=True)
    for name, m in MODELS.items():
        with open(os.path.join(OUT_JSON, name + ".json"), "w") as fh:
            json.dump(m, fh, indent=1)
            fh.write("\n")
    emit_inc(OUT_INC)
    print("wrote", len(MODELS), "models:", ", ".join(MODELS))


if __name__ == "__main__":
    main()
