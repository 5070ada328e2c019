"""N4: model-minus-reference statistics on device -- the immediate consumer of the TBs.

`stats_by_channel` mirrors `/root/reference/python_src/plot_scripts/multi_campaign_plots_and_ana.py:155-178`
(same name, argument meaning and return order: std, bias, rmse) but evaluates every trailing column
(channel x elevation x crop ...) in one launch of `k_stats` instead of a Python loop per channel.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .forward import _need_cuda, _ptr, _stream


def stats_by_channel(values, references, n_chans=None):
    """values, references: (time, ...) arrays or CUDA tensors (NaN = missing).  Returns
    (std, bias, rmse) with the shape of the trailing dimensions, as CUDA fp64 tensors.
    n_valid counts non-NaN `values` (the reference's choice), NaN deviations are skipped."""
    _need_cuda()
    lib = _lib.load()

    def dev(x):
        if not isinstance(x, torch.Tensor):
            x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64))
        return x.to("cuda", torch.float64).contiguous()
    v, r = dev(values), dev(references)
    if v.shape != r.shape:
        raise ValueError("values and references must have the same shape")
    if n_chans is not None and v.shape[1] != n_chans:
        raise ValueError("n_chans does not match axis 1")
    T = v.shape[0]
    tail = tuple(v.shape[1:])
    M = int(np.prod(tail)) if tail else 1
    bias = torch.empty(tail, dtype=torch.float64, device="cuda")
    sd, rmse, nv = torch.empty_like(bias), torch.empty_like(bias), torch.empty_like(bias)
    _lib.check(lib.mwrt_stats_f64(int(T), M, _ptr(v), _ptr(r), _ptr(bias), _ptr(sd), _ptr(rmse), _ptr(nv),
                                  _stream()), "stats")
    torch.cuda.current_stream().synchronize()
    return sd, bias, rmse


def deviations(model_tb, reference_tb):
    """`Deviations_<var>_<ref>` of x_analysis_script_MARCH26.py:181-230: plain difference, NaN-propagating."""
    return model_tb - reference_tb
