"""Training losses on the K4 fused masked-reduction kernels (csrc/loss.cu).

Same names, keyword arguments and values as Hy2DL/aux_functions/functions_training.py:4-83.  The
optional `group` argument is new: under data parallelism the masked partial sums are all-reduced
before the mean is formed, so that N ranks with batch B each give exactly the loss (and gradient)
of one process with batch N*B even when the NaN counts differ between ranks.
"""
from __future__ import annotations

import torch

from ..ops import NseLossFunction, WrmseLossFunction


def _same_order(*tensors):
    """Bring all arguments to one contiguous element order (kernels are order-agnostic)."""
    shape = tensors[0].shape
    out = []
    for t in tensors:
        if t.shape != shape:
            t = t.reshape(shape) if t.numel() == tensors[0].numel() else t.expand(shape)
        out.append(t.contiguous())
    return out


def nse_basin_averaged(y_sim: torch.Tensor, y_obs: torch.Tensor, per_basin_target_std: torch.Tensor,
                       group=None) -> torch.Tensor:
    """Basin-averaged NSE loss: mean over non-NaN targets of (y_sim - y_obs)^2 / (std + 0.1)^2."""
    y_sim, y_obs, std = _same_order(y_sim, y_obs, per_basin_target_std)
    return NseLossFunction.apply(y_sim, y_obs, std, group)


def weighted_rmse(y_sim: torch.Tensor, y_obs: torch.Tensor, group=None) -> torch.Tensor:
    """0.75 RMSE(y) + 0.25 RMSE(log10(sqrt(y + 1e-6) + 0.1)) over non-NaN targets."""
    y_sim, y_obs = _same_order(y_sim, y_obs)
    return WrmseLossFunction.apply(y_sim, y_obs, group)
