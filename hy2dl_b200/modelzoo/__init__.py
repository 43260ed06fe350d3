"""Drop-in counterparts of Hy2DL/modelzoo (same class names, constructor arguments, state-dict keys
and return dictionaries), running on the hy2dl_b200 CUDA kernels."""
from .baseconceptualmodel import BaseConceptualModel
from .cudalstm import CudaLSTM
from .hbv import HBV
from .hybrid import Hybrid
from .linear_reservoir import linear_reservoir
from .lstm import LSTM
from .mflstm import MFLSTM
from .nonsense import NonSense
from .shm import SHM
from .uh_routing import UH_routing

__all__ = ["BaseConceptualModel", "CudaLSTM", "HBV", "Hybrid", "LSTM", "MFLSTM", "NonSense", "SHM", "UH_routing", "linear_reservoir"]
