"""CPU check of the table-driven exp/log the kernels are built from (csrc/mwrt_common.cuh): the tables
and constants are parsed out of the CUDA header and the algorithms are replayed in NumPy (fused
multiply-adds emulated in extended precision), so a wrong table digit, constant or polynomial degree
is caught without a GPU.  The GPU test of the real device code is
tests/test_gpu_parity.py::test_device_math_helpers."""
import os
import re

import numpy as np

HDR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                   "mwr_fast_forward_operators_and_lbls_b200", "csrc", "mwrt_common.cuh")
LD = np.longdouble


def _table(name):
    src = open(HDR).read()
    m = re.search(r"%s\[[^\]]*\]\s*=\s*\{(.*?)\};" % re.escape(name), src, flags=re.S)
    assert m, name
    return np.array([float.fromhex(t) for t in re.findall(r"-?0x[0-9a-fA-F.]+p[+-]?\d+", m.group(1))])


def _const(name):
    src = open(HDR).read()
    m = re.search(r"constexpr double %s\s*=\s*(-?0x[0-9a-fA-F.]+p[+-]?\d+)" % re.escape(name), src)
    assert m, name
    return float.fromhex(m.group(1))


def fma(a, b, c):
    return np.asarray(LD(a) * LD(b) + LD(c), dtype=np.float64)


def test_exp_table_and_algorithm():
    tab = _table("kExpTab")
    assert tab.size == 256
    j = np.arange(256)
    assert np.max(np.abs(tab - 2.0 ** (j / 256.0)) / tab) < 2.3e-16          # correctly rounded entries
    inv_step, step = _const("kExpInvStep"), _const("kExpStep")
    assert abs(inv_step * np.log(2.0) / 256.0 - 1.0) < 3e-16 and abs(step * 256.0 / np.log(2.0) - 1.0) < 3e-16
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.uniform(-700, 700, 100000), rng.uniform(-2, 2, 100000), [0.0, 1e-300, -1e-9]])
    magic = 6755399441055744.0
    t = fma(x, inv_step, magic)
    n = (t - magic).astype(np.int64)           # what the low mantissa bits hold
    r = fma(-(t - magic), step, x)
    T = tab[n & 255]
    p = fma(r, 4.1666666666666664e-02, 1.6666666666666666e-01)
    p = fma(p, r, 0.5)
    p = fma(p, r, 1.0)
    res = fma(T, p * r, T) * 2.0 ** (n >> 8).astype(np.float64)
    ref = np.exp(x)
    rel = np.abs(res - ref) / ref
    # (the extended-precision stand-in for the fma can mis-round exact ties of the reduction by 5e-4 steps)
    assert np.all(np.abs(r) <= np.log(2.0) / 512 * (1 + 2e-3))
    assert np.max(rel / (2.3e-16 + 1.2e-16 * np.abs(x))) <= 1.0


def test_log_table_and_algorithm():
    tab = _table("kLogTab64").reshape(64, 2)
    c, lg = tab[:, 0], tab[:, 1]
    assert c[0] == 1.0 and lg[0] == 0.0 and c[63] == 1.0 and lg[63] == 0.0    # exact around x = 1
    assert np.max(np.abs(lg + np.log(c))) < 1.2e-16                           # -log(c) correctly rounded
    rng = np.random.default_rng(6)
    x = np.concatenate([np.exp(rng.uniform(-600, 600, 100000)), rng.uniform(0.5, 2.0, 100000),
                        1.0 + rng.uniform(-1e-6, 1e-6, 20000), 1.0 + rng.uniform(-0.03, 0.03, 20000),
                        [1.0, 2.0, 0.5, 1.0 + 27 / 64, np.nextafter(1.0 + 27 / 64, 0)]])
    m, e = np.frexp(x)                         # x = m * 2^e, m in [0.5, 1)
    m, e = m * 2.0, e - 1                      # m in [1, 2)
    i = np.minimum(((m - 1.0) * 64).astype(np.int64), 63)
    big = i >= 27
    m = np.where(big, m * 0.5, m)
    e = np.where(big, e + 1, e).astype(np.float64)
    r = fma(m, c[i], -1.0)
    assert np.max(np.abs(r)) <= 1.0 / 64 + 1e-12
    q = np.full_like(r, 1.0 / 9.0)
    for coef in (-1.0 / 8.0, 1.0 / 7.0, -1.0 / 6.0, 1.0 / 5.0, -1.0 / 4.0, 1.0 / 3.0, -0.5):
        q = fma(q, r, coef)
    tail = fma(r * r, q, e * 1.90821492927058770002e-10)
    res = fma(e, 6.93147180369123816490e-01, lg[i]) + (r + tail)
    ref = np.log(x)
    nz = ref != 0.0
    assert np.all(res[~nz] == 0.0)
    ulps = np.abs(res[nz] - ref[nz]) / np.spacing(np.abs(ref[nz]))
    assert ulps.max() <= 2.5
