"""Linear reservoir (drop-in for Hy2DL/modelzoo/linear_reservoir.py): one storage filled by precipitation,
emptied by scaled evapotranspiration and a linear outflow, stepped per series by the K3 CUDA kernel."""
from __future__ import annotations

from typing import Dict, List, Tuple

from .. import _lib
from .baseconceptualmodel import BaseConceptualModel


class linear_reservoir(BaseConceptualModel):
    """Parameters ki, aux_ET (linear_reservoir.py:105-108); reads the precipitation and evapotranspiration
    columns only (linear_reservoir.py:79-80), so x_conceptual may have 2, 3 or 4 columns."""

    _model_id = _lib.MODEL_LINRES
    _forcing_columns = 2

    def __init__(self, n_models: int = 1, parameter_type: List[str] = None):
        super().__init__()
        self.n_conceptual_models = n_models
        self.parameter_type = self._map_parameter_type(parameter_type=parameter_type)
        self.output_size = 1

    @property
    def _initial_states(self) -> Dict[str, float]:
        return {"si": 0.001}

    @property
    def parameter_ranges(self) -> Dict[str, Tuple[float, float]]:
        return {"ki": (0.002, 1.0), "aux_ET": (0.0, 1.5)}
