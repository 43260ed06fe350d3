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
HYBRID_CASES = ["hybrid_shm_c2", "hybrid_shm_mixed", "hybrid_hbv_h64", "hybrid_hbv_norouting", "hybrid_nonsense",
                "hybrid_linres", "hybrid_hbv_c4", "hybrid_hbv_c4_t730"]
CUDALSTM_CASES = ["cudalstm_c1", "cudalstm_c3", "cudalstm_c3_full", "cudalstm_c3_h128", "cudalstm_c3_trained",
                  "cudalstm_h100"]


def cudalstm_weights_np(g):
    """numpy state dict of a cudalstm_* golden file (meta = B, T, I, H, wseed[, trained])."""
    meta = [int(v) for v in g["meta"]]
    B, T, I, H, wseed = meta[:5]
    w = lstm_weights(I, H, 1, seed=wseed)
    if len(meta) > 5 and meta[5]:            # "trained-like" weights: make_golden.trained_like
        for k in ("lstm.weight_ih_l0", "lstm.weight_hh_l0"):
            w[k] = (w[k] * 1.5).astype(np.float32)
    return w


def bucket_of(name):
    """Oracle model key of a *_direct golden file."""
    return {"shm": "shm", "hbv": "hbv", "linres": "linear_reservoir", "nonsense": "nonsense"}[name.split("_")[0]]


def camels_like_inputs(B, T, I, seed, n_dynamic=None):
    """LSTM inputs shaped like the reference's samples (basedataset.py:174-177): a few standardised dynamic forcings
    that vary in time and static attributes that are constant over the window.  (White noise on all I inputs makes the
    H = 256 recurrence amplify rounding errors until fp32 itself is 4e-4 away from fp64 at T = 730.)"""
    from hy2dl_b200.synthetic import make_basins
    nd = n_dynamic if n_dynamic is not None else {25: 3, 31: 5, 41: 6}.get(I, min(5, I))
    data = make_basins("us", n_basins=16, n_rows=T + 30, seed=seed)
    rng = np.random.default_rng(seed)
    bi = rng.integers(0, 16, B)
    e = rng.integers(T - 1, T + 30, B)
    xd = np.stack([data.x_d[b, i - T + 1:i + 1, :nd] for b, i in zip(bi, e)]).astype(np.float64)
    xd = (xd - xd.mean((0, 1))) / (xd.std((0, 1)) + 1e-9)
    xs = np.repeat(rng.normal(0, 1, (B, 1, I - nd)), T, axis=1)
    return np.concatenate([xd, xs], axis=2).astype(np.float32)


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


# Every comparison is also recorded element-wise (tests/conftest.py writes the list to
# gpurun_out/parity_report.json at the end of the session): `frac_outside` is the share of elements with
# |a-b| > rtol*|b| + 1e-6*max|b|, i.e. outside numpy.allclose(rtol, atol = 1e-6 * scale) -- the figure a
# tensor-level bound hides for small-magnitude elements (low flows, tiny gradients).
PARITY_LOG = []
_current_test = [""]
# hard element-wise gate on top of the tensor-level one: at most this share of the elements may be outside
# |a-b| <= 10*rtol*|b| + 1e-5*max|b|  (a value that is 1e-5 of the tensor's scale is below the fp32 noise of
# the reference's own recurrence; see test_oracle_golden's fp32-vs-fp64 spread)
ELEMENTWISE_MAX_FRAC = 1e-3


def gate_fraction(a, b, rtol, atol=0.0):
    """share of the elements outside the element-wise gate |a-b| <= 10*rtol*|b| + 1e-5*max|b| + atol"""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    d = np.abs(a - b)
    ok = ~np.isnan(d)
    return float((d[ok] > 10 * rtol * np.abs(b[ok]) + 1e-5 * np.nanmax(np.abs(b)) + atol).sum()) / max(int(ok.sum()), 1)


def assert_close(a, b, rtol, atol=0.0, what="", max_frac=None):
    """max |a-b| <= atol + rtol * max|b|  (tensor-level relative tolerance; stated in each test),
    plus the element-wise gate above (share allowed outside: ELEMENTWISE_MAX_FRAC unless the test calibrates
    `max_frac` itself and says how); both figures are logged."""
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    if isinstance(b, torch.Tensor):
        b = b.detach().cpu().numpy()
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    assert np.isnan(a).sum() == np.isnan(b).sum(), f"{what}: NaN pattern differs"
    if not a.size:
        return
    scale = np.nanmax(np.abs(b))
    d = np.abs(a - b)
    err = np.nanmax(d)
    lim = atol + rtol * scale
    ok = ~np.isnan(d)
    n = max(int(ok.sum()), 1)
    frac_outside = float((d[ok] > rtol * np.abs(b[ok]) + 1e-6 * scale).sum()) / n
    frac_gate = float((d[ok] > 10 * rtol * np.abs(b[ok]) + 1e-5 * scale + atol).sum()) / n
    PARITY_LOG.append({"test": _current_test[0], "what": what, "n": int(a.size), "rtol": rtol, "scale": float(scale),
                       "max_err_over_scale": float(err / (scale + 1e-300)), "frac_outside_elementwise": frac_outside,
                       "frac_outside_gate": frac_gate})
    assert err <= lim, f"{what}: max abs err {err:.3e} > {lim:.3e} (rtol {rtol}, scale {scale:.3e})"
    assert frac_gate * n <= max(1.0, (max_frac or ELEMENTWISE_MAX_FRAC) * n), (f"{what}: {frac_gate:.2e} of the {n} elements are outside "
                                                                 f"|a-b| <= {10 * rtol:g}|b| + 1e-5*scale")
