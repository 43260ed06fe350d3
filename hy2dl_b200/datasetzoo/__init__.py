"""GPU-resident counterpart of Hy2DL/datasetzoo's sampler (BaseDataset.__getitem__ + DataLoader)."""
from .resident import Batch, ResidentDataset, valid_window_flags

__all__ = ["Batch", "ResidentDataset", "valid_window_flags"]
