"""HBV bucket model (drop-in for Hy2DL/modelzoo/hbv.py): snowpack, meltwater, soil moisture, upper
and lower zone stepped per series by the K3 CUDA kernel (csrc/conceptual.cu)."""
from __future__ import annotations

from typing import Dict, List, Tuple

from .. import _lib
from .baseconceptualmodel import BaseConceptualModel


class HBV(BaseConceptualModel):
    """13 parameters incl. BETAET (hbv.py:199-215); a parameter dict without "BETAET" selects the
    linear evaporation factor (hbv.py:157-162)."""

    _model_id = _lib.MODEL_HBV

    def __init__(self, n_models: int = 1, parameter_type: List[str] = None):
        super().__init__()
        self.n_conceptual_models = n_models
        self.parameter_type = self._map_parameter_type(parameter_type=parameter_type)
        self.output_size = 1

    @property
    def _initial_states(self) -> Dict[str, float]:
        return {"SNOWPACK": 0.001, "MELTWATER": 0.001, "SM": 0.001, "SUZ": 0.001, "SLZ": 0.001}

    @property
    def parameter_ranges(self) -> Dict[str, Tuple[float, float]]:
        return {"BETA": (1.0, 6.0), "FC": (50.0, 1000.0), "K0": (0.05, 0.9), "K1": (0.01, 0.5), "K2": (0.001, 0.2),
                "LP": (0.2, 1.0), "PERC": (0.0, 10.0), "UZL": (0.0, 100.0), "TT": (-2.5, 2.5),
                "CFMAX": (0.5, 10.0), "CFR": (0.0, 0.1), "CWH": (0.0, 0.2), "BETAET": (0.3, 5.0)}
