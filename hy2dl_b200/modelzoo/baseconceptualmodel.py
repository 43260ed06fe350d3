"""Contract shared by the conceptual (bucket) models -- counterpart of
Hy2DL/modelzoo/baseconceptualmodel.py:7-161 -- expressed on top of the K3 kernels."""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple, Union

import torch
import torch.nn as nn

from .. import _lib, ops


class BaseConceptualModel(nn.Module):
    """Abstract base: subclasses provide `parameter_ranges`, `_initial_states` and `_model_id`."""

    _model_id: Optional[int] = None  # HY2DL_MODEL_* of the CUDA implementation
    _forcing_columns: Optional[int] = None  # leading columns of x_conceptual the model reads (None: all, 3 or 4)

    def __init__(self):
        super().__init__()

    # ---- reference API ------------------------------------------------------------------------
    def forward(self, x_conceptual: torch.Tensor, parameters: Dict[str, torch.Tensor],
                initial_states: Optional[Dict[str, torch.Tensor]] = None
                ) -> Dict[str, Union[torch.Tensor, Dict[str, torch.Tensor]]]:
        """x_conceptual [B,T,3|4], parameters {name: [B,T,m]} (time-expanded views count as static),
        initial_states {name: [B,m]} -> the reference's result dictionary (shm.py:187, hbv.py:193)."""
        if self._model_id is None:
            raise NotImplementedError
        B, T, _ = x_conceptual.shape
        m = self.n_conceptual_models
        names = list(self.parameter_ranges)
        planes, offs, dynamic, off = [], [], [], 0
        for n in names:
            if n not in parameters:
                if not (self._model_id == _lib.MODEL_HBV and n == "BETAET"):
                    raise KeyError(f"parameter '{n}' missing")
                offs.append(None)
                dynamic.append(False)
                continue
            p = parameters[n]
            if tuple(p.shape) != (B, T, m):
                raise ValueError(f"parameter '{n}' has shape {tuple(p.shape)}, expected {(B, T, m)}")
            is_dyn = not (p.stride(1) == 0 or T == 1)
            plane = p.permute(1, 0, 2).reshape(T, B * m) if is_dyn else p[:, 0, :].reshape(B * m)
            pad = (-plane.numel()) % 4
            planes.append(plane.reshape(-1))
            if pad:
                planes.append(plane.new_zeros(pad))
            offs.append(off)
            dynamic.append(is_dyn)
            off += plane.numel() + pad
        par_flat = torch.cat(planes)
        init = None
        if initial_states is not None:
            init = torch.stack([initial_states[k].reshape(B * m) for k in self._initial_states])
        forc = ops.pack_forcing(self._forcing(x_conceptual))
        states, q = ops.conceptual(forc, par_flat, init, self._model_id, B, m, 0, T, offs, dynamic)
        return self._result(states, q, parameters, B, m)

    def _forcing(self, x_conceptual: torch.Tensor) -> torch.Tensor:
        """The columns of x_conceptual the model reads (SHM/HBV: all three or four)."""
        n = self._forcing_columns
        if n is None:
            return x_conceptual
        if x_conceptual.shape[2] < n:
            raise ValueError(f"{type(self).__name__} needs {n} conceptual input columns, got {x_conceptual.shape[2]}")
        return x_conceptual[:, :, :n]

    def _result(self, states, q, parameters, B: int, m: int):
        T = q.shape[0]
        internal = {k: states[:, i, :].view(T, B, m).permute(1, 0, 2) for i, k in enumerate(self._initial_states)}
        final = {k: states[T - 1, i, :].view(B, m) for i, k in enumerate(self._initial_states)}
        return {"y_hat": q.t().unsqueeze(2), "parameters": parameters, "internal_states": internal,
                "final_states": final}

    def map_parameters(self, lstm_out: torch.Tensor, warmup_period: int
                       ) -> Tuple[Dict[str, torch.Tensor], Dict[str, torch.Tensor]]:
        """Reference-compatible mapping of a [B,T,P*m] head output to (warm-up, simulation) dicts
        (baseconceptualmodel.py:22-78).  The Hybrid class does not call this -- its head kernel
        fuses Linear + mapping -- it is kept for users who call it directly; plain tensor ops."""
        lstm_out = lstm_out.view(lstm_out.shape[0], lstm_out.shape[1], -1, self.n_conceptual_models)
        T = lstm_out.shape[1]
        warm, sim = {}, {}
        for index, (name, (lo, hi)) in enumerate(self.parameter_ranges.items()):
            kind = self.parameter_type[name]
            if kind == "static":
                zw = lstm_out[:, -1:, index, :].expand(-1, warmup_period, -1)
                zs = lstm_out[:, -1:, index, :].expand(-1, T - warmup_period, -1)
            elif kind == "dynamic":
                zw = lstm_out[:, warmup_period - 1:warmup_period, index, :].expand(-1, warmup_period, -1)
                zs = lstm_out[:, warmup_period:, index, :]
            else:
                raise ValueError(f"Unsupported parameter type {kind}")
            warm[name] = lo + torch.sigmoid(zw) * (hi - lo)
            sim[name] = lo + torch.sigmoid(zs) * (hi - lo)
        return warm, sim

    def _map_parameter_type(self, parameter_type: List[str] = None) -> Dict[str, str]:
        chosen = set(parameter_type or ())
        return {key: ("dynamic" if key in chosen else "static") for key in self.parameter_ranges}

    def _get_final_states(self, states: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        return {name: state[:, -1, :] for name, state in states.items()}

    @property
    def _initial_states(self) -> Dict[str, float]:
        raise NotImplementedError

    @property
    def parameter_ranges(self) -> Dict[str, Tuple[float, float]]:
        raise NotImplementedError
