"""NonSense bucket model (drop-in for Hy2DL/modelzoo/nonsense.py): water passes snow -> baseflow -> interflow ->
unsaturated zone before it leaves, a physically implausible ordering used as a control in the hybrid-model study.
Stepped per series by the K3 CUDA kernel (csrc/conceptual.cu)."""
from __future__ import annotations

from typing import Dict, List, Tuple

from .. import _lib
from .baseconceptualmodel import BaseConceptualModel


class NonSense(BaseConceptualModel):
    """Parameters dd, sumax, beta, ki, kb (nonsense.py:174-176); storages ss, su, si, sb with defaults 0, 5, 10, 15
    (nonsense.py:170-172).  Temperature is column 2 of x_conceptual, whatever the number of columns
    (nonsense.py:80).  The reference writes the [B, n_models] outflow into a [B, 1] slot (nonsense.py:163),
    which only works for n_models = 1; more members are rejected here as they fail there."""

    _model_id = _lib.MODEL_NONSENSE
    _forcing_columns = 3

    def __init__(self, n_models: int = 1, parameter_type: List[str] = None):
        super().__init__()
        if n_models != 1:
            raise ValueError("NonSense supports n_models = 1 only (the reference's output slot is [batch, 1], "
                             "nonsense.py:163)")
        self.n_conceptual_models = n_models
        self.parameter_type = self._map_parameter_type(parameter_type=parameter_type)
        self.output_size = 1

    @property
    def _initial_states(self) -> Dict[str, float]:
        return {"ss": 0.0, "su": 5.0, "si": 10.0, "sb": 15.0}

    @property
    def parameter_ranges(self) -> Dict[str, Tuple[float, float]]:
        return {"dd": (0.0, 10.0), "sumax": (20.0, 700.0), "beta": (1.0, 6.0), "ki": (1.0, 100.0),
                "kb": (10.0, 1000.0)}
