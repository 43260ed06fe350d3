"""Shared helpers for the parity tests: golden loading and oracle drivers (test infrastructure)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import hy2dl_oracle as O  # noqa: E402
from hy2dl_b200.synthetic import lstm_weights  # noqa: E402

PARAM_KEYS = ("lstm.weight_ih_l0", "lstm.weight_hh_l0", "lstm.bias_ih_l0", "lstm.bias_hh_l0",
              "linear.weight", "linear.bias")


def golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


DIRECT_CASES = ["shm_direct", "shm_direct_static4", "hbv_direct", "hbv_direct_nobetaet", "linres_direct",
                "linres_direct_static", "nonsense_direct", "nonsense_direct_static"]
HYBRID_CASES = ["hybrid_shm_c2", "hybrid_shm_mixed", "hybrid_hbv_c4", "hybrid_hbv_norouting", "hybrid_nonsense",
                "hybrid_linres"]


def bucket_of(name):
    """Oracle model key of a *_direct golden file."""
    return {"shm": "shm", "hbv": "hbv", "linres": "linear_reservoir", "nonsense": "nonsense"}[name.split("_")[0]]


def tt(a, dtype=torch.float32, grad=False):
    x = torch.tensor(np.asarray(a), dtype=dtype)
    return x.requires_grad_(grad)


def weights(I, H, n_out, seed, dtype=torch.float32, grad=True):
    return {k: tt(v, dtype, grad) for k, v in lstm_weights(I, H, n_out, seed=seed).items()}


def oracle_loss(kind, y, g, dtype):
    if kind == "nse":
        return O.nse_basin_averaged(y, tt(g["y_obs"], dtype), tt(g["basin_std"], dtype))
    return O.weighted_rmse(y, tt(g["y_obs"], dtype))


def oracle_hybrid(g, dtype=torch.float32, library=False):
    B, T, n_last, I, H, m, wseed, routing = [int(v) for v in g["meta"]]
    model = str(g["model"])
    dynamic = [str(s) for s in g["dynamic"]]
    n_out = len(O.MODEL_TABLE[model][0]) * m + (2 if routing else 0)
    w = weights(I, H, n_out, wseed, dtype)
    pred = O.hybrid_forward(w, tt(g["x_lstm"], dtype), tt(g["x_conceptual"], dtype), model=model, n_models=m,
                            dynamic=dynamic, warmup=T - n_last, routing=bool(routing), library=library)
    loss = oracle_loss(str(g["loss_kind"]), pred["y_hat"], g, dtype)
    loss.backward()
    return w, pred, loss


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-30))


def assert_close(a, b, rtol, atol=0.0, what=""):
    """max |a-b| <= atol + rtol * max|b|  (tensor-level relative tolerance; stated in each test)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    assert np.isnan(a).sum() == np.isnan(b).sum(), f"{what}: NaN pattern differs"
    err = np.nanmax(np.abs(a - b)) if a.size else 0.0
    lim = atol + rtol * (np.nanmax(np.abs(b)) if b.size else 0.0)
    assert err <= lim, f"{what}: max abs err {err:.3e} > {lim:.3e} (rtol {rtol}, scale {np.nanmax(np.abs(b)):.3e})"
