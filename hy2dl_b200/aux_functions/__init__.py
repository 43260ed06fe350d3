"""Drop-in counterparts of Hy2DL/aux_functions for the hot path (losses) plus the evaluation metric."""
from .functions_evaluation import nse, nse_device
from .functions_training import nse_basin_averaged, weighted_rmse

__all__ = ["nse", "nse_device", "nse_basin_averaged", "weighted_rmse"]
