"""hy2dl_b200 -- B200-native (sm_100a) implementation of Hy2DL's training/inference hot path.

Public surface mirrors the reference's module layout:
    hy2dl_b200.modelzoo       CudaLSTM, Hybrid, SHM, HBV, UH_routing, LSTM
    hy2dl_b200.aux_functions  nse_basin_averaged, weighted_rmse, nse
    hy2dl_b200.datasetzoo     ResidentDataset (GPU-resident sliding-window sampler)
    hy2dl_b200.training       FusedClipAdam, Trainer (notebook train-step body)
Everything that computes runs in libhy2dl_b200.so (hy2dl_b200/csrc, C ABI in include/hy2dl_b200.h);
there is no CPU / eager-PyTorch fallback -- a missing library or a non-CUDA tensor raises.
"""
import sys as _sys

__version__ = "0.1.0"


def install_flat_aliases() -> None:
    """Make the reference notebooks' flat imports resolve to this package.

    The notebooks do `sys.path.append("../modelzoo")` and then `from hybrid import Hybrid`,
    `from shm import SHM`, `from functions_training import nse_basin_averaged` ...
    (experiments/HybridModel_GB.ipynb:53-66).  After this call those statements import the
    hy2dl_b200 classes instead, with no other change to the notebook.
    """
    from . import aux_functions, modelzoo
    from .aux_functions import functions_evaluation, functions_training
    from .modelzoo import baseconceptualmodel, cudalstm, hbv, hybrid, nonsense, shm, uh_routing
    # the package attribute `modelzoo.linear_reservoir` is the class (same name as its module, as in the reference)
    linear_reservoir = _sys.modules[__name__ + ".modelzoo.linear_reservoir"]
    for name, mod in {
        "cudalstm": cudalstm, "hybrid": hybrid, "shm": shm, "hbv": hbv, "uh_routing": uh_routing,
        "baseconceptualmodel": baseconceptualmodel, "linear_reservoir": linear_reservoir, "nonsense": nonsense,
        "functions_training": functions_training, "functions_evaluation": functions_evaluation,
    }.items():
        _sys.modules[name] = mod
