"""GPU-resident sliding-window sampler.

The reference keeps every basin as CPU tensors and builds each sample in Python
(datasetzoo/basedataset.py:168-195), then collates 256 of them and copies the batch to the device
every step -- measured at ~11 k sequences/s, the real bottleneck of the reference on any GPU
(SURVEY.md section 6).  Here the standardised series of all basins live in HBM once (~70 MB for
CAMELS-GB) and one kernel launch (csrc/sampler.cu) gathers a batch of (basin, end index) windows
directly into the time-major layouts the LSTM and bucket kernels read.

Validity rules and statistics follow basedataset.py:216-267 (basin std, global scaler) and
:311-334 (window validity).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Iterator, Optional, Sequence

import numpy as np
import torch

from .. import _lib
from .._lib import call, ptr, stream
from ..ops import round_up


def valid_window_flags(x: np.ndarray, y: np.ndarray, seq_length: int, predict_last_n: int,
                       attributes: Optional[np.ndarray] = None, check_nan: bool = True) -> np.ndarray:
    """bool [N]: row i can end a sample (basedataset.py:311-334), vectorised with prefix sums."""
    n = x.shape[0]
    ok = np.zeros(n, dtype=bool)
    if n < seq_length:
        return ok
    ok[seq_length - 1:] = True
    if not check_nan:
        return ok
    if attributes is not None and np.isnan(attributes).any():
        return np.zeros(n, dtype=bool)
    bad_x = np.concatenate([[0], np.cumsum(np.isnan(x).any(axis=1))])
    good_y = np.concatenate([[0], np.cumsum(~np.isnan(y).all(axis=1) if y.ndim > 1 else ~np.isnan(y))])
    i = np.arange(seq_length - 1, n)
    ok[i] &= (bad_x[i + 1] - bad_x[i + 1 - seq_length]) == 0
    ok[i] &= (good_y[i + 1] - good_y[i + 1 - predict_last_n]) > 0
    return ok


@dataclass
class Batch:
    """One gathered batch in native layouts (include/hy2dl_b200.h)."""
    xT: torch.Tensor                 # [L, B, IP], or RI32 [L, G, 1, 32, 32]
    forc: Optional[torch.Tensor]     # [L, 3, B]
    y_obs: torch.Tensor              # [n_last, B]
    basin_std: torch.Tensor          # [n_last, B]
    basin: torch.Tensor              # [B] int64 basin index
    end: torch.Tensor                # [B] int64 end row within the basin
    input_size: int

    @property
    def B(self) -> int:
        return self.basin.numel()

    def as_reference(self) -> Dict[str, torch.Tensor]:
        """The reference's sample dict ([B, T, ...] tensors), for API-level comparisons."""
        from ..ops import from_ri32, is_ri32
        xT = from_ri32(self.xT, self.B) if is_ri32(self.xT) else self.xT
        d = {"x_lstm": xT[:, :, :self.input_size].permute(1, 0, 2).contiguous(),
             "y_obs": self.y_obs.t().unsqueeze(2).contiguous(),
             "basin_std": self.basin_std.t().unsqueeze(2).contiguous()}
        if self.forc is not None:
            d["x_conceptual"] = self.forc.permute(2, 0, 1).contiguous()
        return d


class ResidentDataset:
    """x_d [nb, nr, nd], x_s [nb, ns] or None, x_conc [nb, nr, nc] or None, y_obs [nb, nr, 1] (numpy)."""

    def __init__(self, x_d: np.ndarray, y_obs: np.ndarray, sequence_length: int, predict_last_n: int = 1,
                 x_s: Optional[np.ndarray] = None, x_conc: Optional[np.ndarray] = None, device="cuda",
                 check_nan: bool = True, basins: Optional[Sequence[int]] = None):
        # Data-parallel shard (`basins` = this rank's basins): the windows and the resident series are the shard's,
        # but the scaler is a GLOBAL statistic (basedataset.py:233-245 stacks every basin) -- every rank must
        # standardise with the same numbers or N ranks no longer reproduce the single-process batch.  So the full
        # arrays stay referenced for calculate_global_statistics() and only the shard is uploaded.
        self._full = (x_d, x_s, y_obs)
        self.basins = None if basins is None else np.asarray(basins)
        if basins is not None:
            basins = self.basins
            x_d, y_obs = x_d[basins], y_obs[basins]
            x_s = None if x_s is None else x_s[basins]
            x_conc = None if x_conc is None else x_conc[basins]
        self.sequence_length, self.predict_last_n = sequence_length, predict_last_n
        self.nb, self.nr, self.nd = x_d.shape
        self.ns = 0 if x_s is None else x_s.shape[1]
        self.nc = 0 if x_conc is None else x_conc.shape[2]
        self.input_size = self.nd + self.ns
        self.IP = round_up(self.input_size, 8)
        self.device = torch.device(device)
        self._x_d, self._x_s, self._x_conc, self._y = x_d, x_s, x_conc, y_obs
        inputs = x_d if x_conc is None else np.concatenate([x_d, x_conc], axis=2)
        ve = []
        for b in range(self.nb):
            ok = valid_window_flags(inputs[b], y_obs[b], sequence_length, predict_last_n,
                                    None if x_s is None else x_s[b], check_nan)
            idx = np.nonzero(ok)[0]
            ve.append(np.stack([np.full_like(idx, b), idx], axis=1))
        self.valid_entities = np.concatenate(ve, axis=0).astype(np.int64)
        self.scaler: Dict[str, np.ndarray] = {}
        self.basin_std = np.zeros(self.nb, dtype=np.float32)
        self._resident = False
        self.layout = "rowmajor"   # layout gather() writes xT in; "ri32" for precision tf32x3 models

    def __len__(self) -> int:
        return len(self.valid_entities)

    # ---- statistics (training period only; basedataset.py:216-267) ------------------------------
    def calculate_basin_std(self):
        self.basin_std = np.array([np.nanstd(self._y[b]) for b in range(self.nb)], dtype=np.float32)

    def calculate_global_statistics(self):
        """Global mean / std over ALL basins handed to the constructor (not only this rank's shard)."""
        full_x_d, full_x_s, full_y = self._full
        gx = full_x_d.reshape(-1, self.nd)
        self.scaler["x_d_mean"] = np.nanmean(gx, axis=0).astype(np.float32)
        self.scaler["x_d_std"] = np.nanstd(gx, axis=0).astype(np.float32)
        gy = full_y.reshape(-1, full_y.shape[2])
        self.scaler["y_mean"] = np.nanmean(gy, axis=0).astype(np.float32)
        self.scaler["y_std"] = np.nanstd(gy, axis=0).astype(np.float32)
        if full_x_s is not None:
            self.scaler["x_s_mean"] = full_x_s.mean(axis=0).astype(np.float32)
            self.scaler["x_s_std"] = full_x_s.std(axis=0, ddof=1).astype(np.float32)

    def set_scaler(self, scaler: Dict[str, np.ndarray]):
        """Use a scaler computed elsewhere (validation / test sets reuse the training scaler; a rank that was only
        given its own shard's arrays receives the global one)."""
        self.scaler = {k: np.asarray(v.numpy() if hasattr(v, "numpy") else v, dtype=np.float32) for k, v in scaler.items()}

    # ---- the reference's dataset object as the source --------------------------------------------
    @classmethod
    def from_reference(cls, dataset, device="cuda", basins: Optional[Sequence[int]] = None) -> "ResidentDataset":
        """Build the resident sampler from a constructed reference `BaseDataset` (any datasetzoo subclass;
        datasetzoo/basedataset.py:56-163): the notebooks' dataset cell stays as it is and this is the one added line.

        Takes the per-basin tensors `sequence_data[basin]["x_d" | "y_obs" | "x_conceptual" | "x_s"]` (:147-163) *in
        the state they are in* -- call it after `calculate_basin_std()`, `calculate_global_statistics()` and
        `standardize_data()` exactly as the notebooks do (LSTM_CAMELS_GB.ipynb: training-set cell) -- together with
        the dataset's own `valid_entities` (so sample k here is `dataset[k]` there), `basin_std` and `scaler`, and
        uploads them.  All basins of a BaseDataset share one date range (:129-131), so the series are rectangular.
        `basins`: optional list of positions in `resident.basin_ids` to keep (data-parallel shard).
        """
        ids = [b for b in dataset.entities_ids if b in dataset.sequence_data]
        if not ids:
            raise ValueError("the reference dataset holds no basin with a valid sample")
        if basins is not None:
            ids = [ids[int(k)] for k in basins]
        sd = dataset.sequence_data
        npy = lambda t: t.detach().cpu().numpy().astype(np.float32)
        x_d = np.stack([npy(sd[b]["x_d"]) for b in ids])
        y = np.stack([npy(sd[b]["y_obs"]) for b in ids])
        if y.shape[2] != 1:
            raise NotImplementedError("hy2dl_b200 handles a single target variable, as every reference experiment does")
        x_s = np.stack([npy(sd[b]["x_s"]) for b in ids]) if "x_s" in sd[ids[0]] else None
        x_c = np.stack([npy(sd[b]["x_conceptual"]) for b in ids]) if "x_conceptual" in sd[ids[0]] else None
        self = cls.__new__(cls)
        self.sequence_length, self.predict_last_n = dataset.sequence_length, dataset.predict_last_n
        self.nb, self.nr, self.nd = x_d.shape
        self.ns = 0 if x_s is None else x_s.shape[1]
        self.nc = 0 if x_c is None else x_c.shape[2]
        self.input_size = self.nd + self.ns
        self.IP = round_up(self.input_size, 8)
        self.device = torch.device(device)
        self._x_d, self._x_s, self._x_conc, self._y = x_d, x_s, x_c, y
        self._full = (x_d, x_s, y)
        self.basins = None
        self.basin_ids = list(ids)
        pos = {b: k for k, b in enumerate(ids)}
        self.valid_entities = np.array([(pos[b], i) for b, i in dataset.valid_entities if b in pos],
                                       dtype=np.int64).reshape(-1, 2)
        self.scaler = {}
        if dataset.scaler:
            self.set_scaler(dataset.scaler)
        self.basin_std = np.array([float(dataset.basin_std[b]) if b in dataset.basin_std else 0.0 for b in ids],
                                  dtype=np.float32)
        self._resident = False
        self.layout = "rowmajor"
        if self.device.type == "cuda":
            self.to_device_raw()        # the tensors are already in the state the notebook left them in
        return self

    def standardize_data(self, standardize_output: bool = True):
        """Standardise and move everything to the GPU (call once, after the statistics)."""
        x_d = (self._x_d - self.scaler["x_d_mean"]) / self.scaler["x_d_std"]
        y = self._y
        if standardize_output:
            y = (y - self.scaler["y_mean"]) / self.scaler["y_std"]
        x_s = None if self._x_s is None else (self._x_s - self.scaler["x_s_mean"]) / self.scaler["x_s_std"]
        self._upload(x_d, x_s, self._x_conc, y)

    def to_device_raw(self):
        """Upload without standardisation (e.g. validation sets that reuse a training scaler first)."""
        self._upload(self._x_d, self._x_s, self._x_conc, self._y)

    def _upload(self, x_d, x_s, x_conc, y):
        dev = self.device
        f32 = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32)).to(dev)
        self.d_x_d = f32(x_d.reshape(self.nb * self.nr, self.nd))
        self.d_y = f32(y.reshape(self.nb * self.nr))
        self.d_x_s = None if x_s is None else f32(x_s)
        self.d_x_conc = None if x_conc is None else f32(x_conc.reshape(self.nb * self.nr, self.nc))
        self.d_std = f32(self.basin_std)
        self.d_row0 = (torch.arange(self.nb, dtype=torch.int64) * self.nr).to(dev)
        self.d_valid = torch.as_tensor(self.valid_entities).to(dev)
        self._resident = True

    # ---- gather ---------------------------------------------------------------------------------
    def gather(self, sample_ids: torch.Tensor) -> Batch:
        """sample_ids: int64 device tensor of indices into valid_entities."""
        if not self._resident:
            raise RuntimeError("call standardize_data() or to_device_raw() first")
        sel = self.d_valid.index_select(0, sample_ids)
        return self.gather_pairs(sel[:, 0].contiguous(), sel[:, 1].contiguous())

    def gather_pairs(self, basin: torch.Tensor, end: torch.Tensor) -> Batch:
        _lib.require_cuda(self.d_x_d, "resident series")
        B, L, n_last, dev = basin.numel(), self.sequence_length, self.predict_last_n, self.device
        ri = self.layout == "ri32"
        if ri and self.input_size > 31:
            raise RuntimeError(f"hy2dl_b200: the RI32 / tf32x3 path takes input_size <= 31 (got {self.input_size})")
        IP = 32 if ri else self.IP
        xT = (torch.empty(L, (B + 31) // 32, 1, 32, 32, dtype=torch.float32, device=dev) if ri
              else torch.empty(L, B, IP, dtype=torch.float32, device=dev))
        forc = torch.empty(L, 3, B, dtype=torch.float32, device=dev) if self.nc else None
        y_win = torch.empty(n_last, B, dtype=torch.float32, device=dev)
        std_win = torch.empty(n_last, B, dtype=torch.float32, device=dev)
        call("hy2dl_window_gather", ptr(self.d_x_d), ptr(self.d_x_s), ptr(self.d_x_conc), ptr(self.d_y),
             ptr(self.d_std), ptr(self.d_row0), ptr(basin), ptr(end), B, L, n_last, self.nd, self.ns, self.nc, IP,
             ptr(xT), ptr(forc), ptr(y_win), ptr(std_win), _lib.LAYOUT_RI32 if ri else _lib.LAYOUT_ROWMAJOR, stream())
        return Batch(xT, forc, y_win, std_win, basin, end, self.input_size)

    def epoch(self, batch_size: int, generator: Optional[torch.Generator] = None, drop_last: bool = True,
              shuffle: bool = True) -> Iterator[Batch]:
        """Shuffled pass over all valid windows (DataLoader(shuffle=True, drop_last=True) semantics)."""
        n = len(self)
        order = (torch.randperm(n, device=self.device, generator=generator) if shuffle
                 else torch.arange(n, device=self.device))
        stop = n - n % batch_size if drop_last else n
        for i in range(0, stop, batch_size):
            yield self.gather(order[i:i + batch_size])
