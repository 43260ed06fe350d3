"""Gamma unit-hydrograph routing (drop-in for Hy2DL/modelzoo/uh_routing.py), K3b CUDA kernels."""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch

from .. import ops
from .baseconceptualmodel import BaseConceptualModel


class UH_routing(BaseConceptualModel):
    """15-tap normalised gamma unit hydrograph, causal convolution with the discharge series."""

    def __init__(self, n_models: int = 1, parameter_type: List[str] = None):
        super().__init__()
        self.n_conceptual_models = 1
        self.parameter_type = self._map_parameter_type()
        self.output_size = 1

    def forward(self, discharge: torch.Tensor, parameters: Dict[str, torch.Tensor]) -> torch.Tensor:
        """discharge [B,T,1]; parameters {"alpha","beta"}: [B,T,1], value at [:,0,0] used (uh_routing.py:42)."""
        q = discharge[:, :, 0].t().contiguous()
        y = ops.uh_route(q, parameters["alpha"][:, 0, 0].contiguous(), parameters["beta"][:, 0, 0].contiguous())
        return y.t().unsqueeze(2)

    @property
    def _initial_states(self) -> Dict[str, float]:
        return {}

    @property
    def parameter_ranges(self) -> Dict[str, Tuple[float, float]]:
        return {"alpha": (0.0, 2.9), "beta": (0.0, 6.5)}
