"""Seeded synthetic CAMELS-shaped basins -- no files, numpy PCG64 only.

Produces, per basin, a continuous daily series with the columns the reference dataset classes hand
to `BaseDataset` (datasetzoo/basedataset.py:109-163): dynamic LSTM inputs, conceptual forcings,
target discharge (with NaNs) and static attributes.  Shapes follow the two dataset families the
benchmark configs name:

  "gb": 3 dynamic (precipitation, peti, temperature) + 22 static, conceptual = the same 3 columns
        (experiments/HybridModel_GB.ipynb:86-115)
  "us": 6 dynamic (prcp, srad, tmax, tmin, vp, dayl) + 35 static for the hybrid, 5 + 26 for the
        plain LSTM; conceptual = (prcp, pet, tmax, tmin)  (experiments/HybridModel_US.ipynb:87-130)
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np

SHAPES = {
    # name: (n_dynamic, n_static, n_conceptual, n_basins, n_rows)
    "gb": (3, 22, 3, 671, 6665),
    "us_lstm": (5, 26, 0, 531, 5842),
    "us": (6, 35, 4, 531, 5842),
}


@dataclass
class SyntheticBasins:
    x_d: np.ndarray          # [n_basins, n_rows, n_dyn]   float32, raw units
    x_s: np.ndarray          # [n_basins, n_static]        float32
    x_conc: Optional[np.ndarray]  # [n_basins, n_rows, n_conc] float32 or None
    y_obs: np.ndarray        # [n_basins, n_rows, 1]       float32, NaN = missing
    shape: str


def make_basins(shape: str = "gb", n_basins: Optional[int] = None, n_rows: Optional[int] = None,
                seed: int = 42, nan_frac: float = 0.02) -> SyntheticBasins:
    nd, ns, nc, nb_default, nr_default = SHAPES[shape]
    nb = n_basins or nb_default
    nr = n_rows or nr_default
    rng = np.random.default_rng(seed)
    doy = (np.arange(nr) % 365).astype(np.float64)

    wet = rng.random((nb, nr)) < 0.6
    prcp = rng.gamma(0.5, 4.0, (nb, nr)) * wet
    pet = np.clip(2.0 + 1.5 * np.sin(2 * np.pi * doy / 365.0) + rng.normal(0, 0.3, (nb, nr)), 0.0, None)
    temp = 8.0 + 10.0 * np.sin(2 * np.pi * (doy - 105.0) / 365.0) + rng.normal(0, 3.0, (nb, nr))

    if shape == "gb":
        x_d = np.stack([prcp, pet, temp], axis=2)
        x_conc = x_d.copy()
    else:
        tmax, tmin = temp + 5.0, temp - 5.0
        srad = 300.0 + 150.0 * np.sin(2 * np.pi * (doy - 80.0) / 365.0) + rng.normal(0, 30.0, (nb, nr))
        vp = 900.0 + 400.0 * np.sin(2 * np.pi * (doy - 110.0) / 365.0) + rng.normal(0, 80.0, (nb, nr))
        dayl = 43200.0 + 12000.0 * np.sin(2 * np.pi * (doy - 80.0) / 365.0) + 0 * prcp
        cols = [prcp, srad, tmax, tmin, vp, dayl][:nd] if nd == 6 else [prcp, srad, tmax, tmin, vp]
        x_d = np.stack(cols, axis=2)
        x_conc = np.stack([prcp, pet, tmax, tmin], axis=2) if nc else None

    y = rng.gamma(1.0, 1.0, (nb, nr, 1))
    y[rng.random((nb, nr, 1)) < nan_frac] = np.nan
    x_s = rng.normal(0, 1, (nb, ns))
    return SyntheticBasins(
        x_d=x_d.astype(np.float32), x_s=x_s.astype(np.float32),
        x_conc=None if x_conc is None else x_conc.astype(np.float32),
        y_obs=y.astype(np.float32), shape=shape,
    )


def lstm_weights(input_size: int, hidden: int, n_out: int, seed: int = 0, forget_bias: float = 3.0):
    """PyTorch-style U(-1/sqrt(H), 1/sqrt(H)) weights from numpy's PCG64 (portable across hosts).

    Keys and shapes are the reference state-dict's (checkpoint row); the forget
    gate bias is set as the notebooks do (experiments/LSTM_CAMELS_GB.ipynb:278).
    """
    rng = np.random.default_rng(seed)
    k = 1.0 / np.sqrt(hidden)
    u = lambda *s: rng.uniform(-k, k, s).astype(np.float32)
    w = {
        "lstm.weight_ih_l0": u(4 * hidden, input_size),
        "lstm.weight_hh_l0": u(4 * hidden, hidden),
        "lstm.bias_ih_l0": u(4 * hidden),
        "lstm.bias_hh_l0": u(4 * hidden),
        "linear.weight": u(n_out, hidden),
        "linear.bias": u(n_out),
    }
    if forget_bias is not None:
        w["lstm.bias_hh_l0"][hidden:2 * hidden] = forget_bias
    return w
