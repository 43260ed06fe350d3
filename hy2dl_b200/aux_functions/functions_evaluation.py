"""Acceptance metric: per-basin Nash-Sutcliffe efficiency in float64, following
Hy2DL/aux_functions/functions_evaluation.py:7-50.  `nse` is the reference's host function (dict of frames);
`nse_device` evaluates the same quantity on series that are already on the GPU (whole-period test loop)."""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch


def nse(df_results: Dict[str, "object"], average: bool = True):
    """df_results: {basin: frame-like with 'y_sim' and 'y_obs' columns}. Median over basins if average."""
    scores = []
    for frame in df_results.values():
        sim = np.asarray(frame["y_sim"], dtype=np.float64)
        obs = np.asarray(frame["y_obs"], dtype=np.float64)
        keep = ~np.isnan(sim)
        sim, obs = sim[keep], obs[keep]
        keep = ~np.isnan(obs)
        sim, obs = sim[keep], obs[keep]
        if sim.size > 1:
            scores.append(1.0 - np.sum((sim - obs) ** 2) / np.sum((obs - obs.mean()) ** 2))
        else:
            scores.append(np.nan)
    return np.nanmedian(scores) if average else np.asarray(scores)


def nse_device(y_sim: torch.Tensor, y_obs: torch.Tensor, average: bool = True) -> torch.Tensor:
    """y_sim, y_obs: [B, T, 1] or [B, T] CUDA tensors (one row per basin, NaN = missing; the `y_hat` a model returns
    fits as is).  Returns the per-basin NSE as a float64 device tensor [B], or its nan-median if `average`
    (functions_evaluation.py:28-50) -- no host round trip of the series."""
    from .. import ops
    from .._lib import call
    ys = y_sim[..., 0] if y_sim.dim() == 3 else y_sim
    yo = y_obs[..., 0] if y_obs.dim() == 3 else y_obs
    if ys.shape != yo.shape or ys.dim() != 2:
        raise ValueError(f"y_sim {tuple(y_sim.shape)} and y_obs {tuple(y_obs.shape)} must both be [B, T(, 1)]")
    ops.require_cuda(ys, "y_sim"); ops.require_cuda(yo, "y_obs")
    B, T = ys.shape
    # time-major buffers: Hybrid's y_hat is a transposed view of one already, so .t().contiguous() is free there
    ys_t, yo_t = ys.detach().t().contiguous(), yo.detach().t().contiguous()
    out = torch.empty(B, dtype=torch.float64, device=ys.device)
    call("hy2dl_nse_basins", ops.ptr(ys_t), ops.ptr(yo_t), T, B, ops.ptr(out), ops.stream())
    return torch.nanmedian(out) if average else out
