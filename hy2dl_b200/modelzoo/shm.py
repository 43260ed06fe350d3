"""SHM bucket model (drop-in for Hy2DL/modelzoo/shm.py): snow, fast-flow, unsaturated-zone,
interflow and baseflow storages stepped per series by the K3 CUDA kernel (csrc/conceptual.cu)."""
from __future__ import annotations

from typing import Dict, List, Tuple

from .. import _lib
from .baseconceptualmodel import BaseConceptualModel


class SHM(BaseConceptualModel):
    """n_models parallel members, averaged; parameter_type lists the dynamic parameters
    (shm.py:28-32).  Constants: shm.py:189-204."""

    _model_id = _lib.MODEL_SHM

    def __init__(self, n_models: int = 1, parameter_type: List[str] = None):
        super().__init__()
        self.n_conceptual_models = n_models
        self.parameter_type = self._map_parameter_type(parameter_type=parameter_type)
        self.output_size = 1

    @property
    def _initial_states(self) -> Dict[str, float]:
        return {"ss": 0.001, "sf": 0.001, "su": 0.001, "si": 0.001, "sb": 0.001}

    @property
    def parameter_ranges(self) -> Dict[str, Tuple[float, float]]:
        return {"dd": (0.0, 10.0), "f_thr": (10.0, 60.0), "sumax": (20.0, 700.0), "beta": (1.0, 6.0),
                "perc": (0.0, 1.0), "kf": (0.05, 0.9), "ki": (0.01, 0.5), "kb": (0.001, 0.2)}
