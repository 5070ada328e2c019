"""CPU restatements of the rows next to the path.  TEST INFRASTRUCTURE.

stats_by_channel        /root/reference/python_src/plot_scripts/multi_campaign_plots_and_ana.py:155-178
derive_TBs4PyRTlib_loop /root/reference/python_src/proc/PyRTlib_processing.py:83-197 (the triple loop,
                        one TbCloudRTE per (time, Crop, elevation) and model, NaN gate included)
"""
import numpy as np

from .pyrtlib_scalar import TbCloudRTE


def stats_by_channel(values, references, n_chans):
    """values, references: (time, n_chans).  Returns (std, bias, rmse) of model minus reference per
    channel with the reference's conventions: NaN deviations are skipped in the sums, but the
    divisor counts the non-NaN MODEL values only (multi_campaign_plots_and_ana.py:170).  Pinned against
    outputs of the reference's own function (tests/golden/make_stats_golden.py)."""
    v = np.asarray(values, dtype=np.float64)[:, :n_chans]
    d = v - np.asarray(references, dtype=np.float64)[:, :n_chans]
    with np.errstate(all="ignore"):
        n = np.array([np.sum(~np.isnan(v[:, c])) for c in range(n_chans)])
        bias = np.array([np.nansum(d[:, c]) for c in range(n_chans)]) / n
        sd = np.sqrt(np.array([np.nansum((d[:, c] - bias[c]) ** 2) for c in range(n_chans)]) / n)
        rmse = np.sqrt(np.array([np.nansum(d[:, c] ** 2) for c in range(n_chans)]) / n)
    return sd, bias, rmse


def check_for_nans(z_in, p_in, t_in, rh_in, frqs, ang):
    return np.any([np.isnan(z_in).any(), np.isnan(p_in).any(), np.isnan(t_in).any(),
                   np.isnan(rh_in).any(), np.isnan(frqs).any(), np.isnan(ang).any()])


def derive_TBs4PyRTlib_loop(level_z, level_p, level_t, level_rh, elevations, frqs,
                            models=("R20", "R24", "R17", "R98")):
    """Level_* arrays (N_Levels, time, Crop), top-first, z [m], RH [%].  Returns dict model ->
    (time, nf, nelev, Crop) exactly as the reference fills tbs/tbs17/tbs98/tbs20."""
    L, T, C = level_z.shape
    nf = len(frqs)
    out = {m: np.full((T, nf, len(elevations), C), np.nan) for m in models}
    for i in range(T):
        for j in range(C):
            for k, elevation in enumerate(elevations):
                ang = np.array([elevation])
                rh_in = level_rh[:, i, j] / 100
                z_in = level_z[:, i, j] / 1000
                p_in = level_p[:, i, j]
                t_in = level_t[:, i, j]
                if check_for_nans(z_in, p_in, t_in, rh_in, frqs, ang):
                    continue
                for mdl in models:
                    rte = TbCloudRTE(z_in[::-1], p_in[::-1], t_in[::-1], rh_in[::-1], frqs, ang)
                    rte.init_absmdl(mdl)
                    rte.satellite = False
                    df = rte.execute()
                    out[mdl][i, :, k, j] = df["tbtotal"]
    return out
