"""CPU oracle for the Hy2DL hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional restatement (torch on the host CPU, dtype-generic so it runs in fp32 for timing
and in fp64 as the arbiter) of the reference algorithm for the path `BASELINE.json:north_star`
names.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module; nothing under `hy2dl_b200/` does.

Reference lines each function follows (paths relative to /root/reference/Hy2DL):
  lstm_cell_sequence ........ torch/nn/modules/rnn.py:842-847 (equations), :935-941 (layout i|f|g|o)
                              call sites modelzoo/cudalstm.py:43-50, modelzoo/hybrid.py:84-92
  cudalstm_forward .......... modelzoo/cudalstm.py:29-55
  map_parameters ............ modelzoo/baseconceptualmodel.py:53-76
  shm_run ................... modelzoo/shm.py:69-93 (pre-loop), :135-182 (step), :189-204 (constants)
  hbv_run ................... modelzoo/hbv.py:70-86 (pre-loop), :127-188 (step), :195-215 (constants)
  linres_run ................ modelzoo/linear_reservoir.py:64-92 (step), :99-108 (constants)
  nonsense_run .............. modelzoo/nonsense.py:72-94 (pre-loop), :128-168 (step), :170-176 (constants)
  mflstm_forward ............ NO REFERENCE CODE: this project's own definition, parity unpinned
  uh_route .................. modelzoo/uh_routing.py:42-44, :64-74, :92-109, :111-113
  hybrid_forward ............ modelzoo/hybrid.py:62-117
  nse_basin_averaged ........ aux_functions/functions_training.py:32-41
  weighted_rmse ............. aux_functions/functions_training.py:72-81
  train_step ................ experiments/LSTM_CAMELS_GB.ipynb:289-299, HybridModel_US.ipynb:280-289
  window_gather ............. datasetzoo/basedataset.py:168-195
  valid_window_flags ........ datasetzoo/basedataset.py:311-334

Third-party arithmetic on the path: `torch.nn.LSTM` (requirements.txt:6 pins torch>=2.2.1; this
image has torch 2.11.0, CPU backend oneDNN).  `lstm_library` calls it (that is what the reference
executes and what the CPU baseline times); `lstm_cell_sequence` restates its published equations.

Parity pin: the reference ships no tests or golden vectors.  This oracle is
pinned against outputs of the *unmodified reference modules imported from /root/reference* on
seeded inputs; the vectors and the script that made them are committed under `tests/golden/`
(`tests/golden/make_golden.py`), and `tests/test_oracle_golden.py` checks every one.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import torch

# ----------------------------------------------------------------------------------------------
# constants of the conceptual models (shm.py:189-204, hbv.py:195-215, uh_routing.py:111-113)
# ----------------------------------------------------------------------------------------------
SHM_RANGES = OrderedDict(
    dd=(0.0, 10.0), f_thr=(10.0, 60.0), sumax=(20.0, 700.0), beta=(1.0, 6.0),
    perc=(0.0, 1.0), kf=(0.05, 0.9), ki=(0.01, 0.5), kb=(0.001, 0.2),
)
SHM_STATES = ("ss", "sf", "su", "si", "sb")
SHM_INIT = 0.001

HBV_RANGES = OrderedDict(
    BETA=(1.0, 6.0), FC=(50.0, 1000.0), K0=(0.05, 0.9), K1=(0.01, 0.5), K2=(0.001, 0.2),
    LP=(0.2, 1.0), PERC=(0.0, 10.0), UZL=(0.0, 100.0), TT=(-2.5, 2.5), CFMAX=(0.5, 10.0),
    CFR=(0.0, 0.1), CWH=(0.0, 0.2), BETAET=(0.3, 5.0),
)
HBV_STATES = ("SNOWPACK", "MELTWATER", "SM", "SUZ", "SLZ")
HBV_INIT = 0.001

LINRES_RANGES = OrderedDict(ki=(0.002, 1.0), aux_ET=(0.0, 1.5))
LINRES_STATES = ("si",)
LINRES_INIT = 0.001

NONSENSE_RANGES = OrderedDict(dd=(0.0, 10.0), sumax=(20.0, 700.0), beta=(1.0, 6.0), ki=(1.0, 100.0), kb=(10.0, 1000.0))
NONSENSE_STATES = ("ss", "su", "si", "sb")
NONSENSE_INIT = {"ss": 0.0, "su": 5.0, "si": 10.0, "sb": 15.0}

UH_RANGES = OrderedDict(alpha=(0.0, 2.9), beta=(0.0, 6.5))
UH_LEN = 15

MODEL_TABLE = {
    "shm": (SHM_RANGES, SHM_STATES, SHM_INIT),
    "hbv": (HBV_RANGES, HBV_STATES, HBV_INIT),
    "linear_reservoir": (LINRES_RANGES, LINRES_STATES, LINRES_INIT),
    "nonsense": (NONSENSE_RANGES, NONSENSE_STATES, NONSENSE_INIT),
}


# ----------------------------------------------------------------------------------------------
# LSTM
# ----------------------------------------------------------------------------------------------
def lstm_cell_sequence(x, w_ih, w_hh, b_ih, b_hh):
    """Explicit cell loop, h0 = c0 = 0, one layer, batch-first. Returns h for every step [B,T,H]."""
    B, T, _ = x.shape
    H = w_hh.shape[1]
    h = x.new_zeros(B, H)
    c = x.new_zeros(B, H)
    gx = x @ w_ih.t() + (b_ih + b_hh)
    hs = []
    for t in range(T):
        a = gx[:, t] + h @ w_hh.t()
        gi, gf, gg, go = a.split(H, dim=1)
        c = torch.sigmoid(gf) * c + torch.sigmoid(gi) * torch.tanh(gg)
        h = torch.sigmoid(go) * torch.tanh(c)
        hs.append(h)
    return torch.stack(hs, dim=1)


def lstm_library(x, w_ih, w_hh, b_ih, b_hh):
    """What the reference executes: torch's LSTM op with zero initial state (cudalstm.py:43-50)."""
    B = x.shape[0]
    H = w_hh.shape[1]
    h0 = x.new_zeros(1, B, H)
    c0 = x.new_zeros(1, B, H)
    # train flag: cuDNN (x on a GPU: bench.py's `baseline_gpu_library` leg) only keeps its reserve space for the backward
    # pass in training mode; with dropout 0 the flag changes no number.  The CPU path (oneDNN) is called as before.
    out, _, _ = torch._VF.lstm(x, (h0, c0), [w_ih, w_hh, b_ih, b_hh], True, 1, 0.0, x.is_cuda, False, True)
    return out


def cudalstm_forward(w: Dict[str, torch.Tensor], x, library: bool = False):
    """Sequence-to-one regressor; dropout is the identity here (eval / p = 0)."""
    f = lstm_library if library else lstm_cell_sequence
    h = f(x, w["lstm.weight_ih_l0"], w["lstm.weight_hh_l0"], w["lstm.bias_ih_l0"], w["lstm.bias_hh_l0"])
    return h[:, -1, :] @ w["linear.weight"].t() + w["linear.bias"]


# ----------------------------------------------------------------------------------------------
# head output -> conceptual parameters
# ----------------------------------------------------------------------------------------------
def mflstm_forward(w: Dict[str, torch.Tensor], x: Dict[str, torch.Tensor], columns: Dict[str, Sequence[int]],
                   flag_column: Dict[str, int], predict_last_n: int):
    """Multi-frequency LSTM -- PARITY UNPINNED: the reference snapshot has no such model, so
    this restates the project's own definition (hy2dl_b200/modelzoo/mflstm.py) in its natural form: one cell,
    stretches processed oldest first, every frequency with its own input projection W_f = W_ih[:, columns[f]] and
    bias offset W_ih[:, flag_column[f]], shared recurrent weights, zero initial state, Linear head on the last steps.
    x = {freq: [B, n_steps, n_inputs]} in processing order."""
    w_ih, w_hh = w["lstm.weight_ih_l0"], w["lstm.weight_hh_l0"]
    bias = w["lstm.bias_ih_l0"] + w["lstm.bias_hh_l0"]
    first = next(iter(x.values()))
    B, H = first.shape[0], w_hh.shape[1]
    h = first.new_zeros(B, H)
    c = first.new_zeros(B, H)
    hs = []
    for f, xf in x.items():
        w_f = w_ih[:, list(columns[f])]
        b_f = bias + (w_ih[:, flag_column[f]] if f in flag_column else 0)
        for t in range(xf.shape[1]):
            z = xf[:, t] @ w_f.t() + h @ w_hh.t() + b_f
            i, fg, g, o = z.chunk(4, dim=1)
            c = torch.sigmoid(fg) * c + torch.sigmoid(i) * torch.tanh(g)
            h = torch.sigmoid(o) * torch.tanh(c)
            hs.append(h)
    out = torch.stack(hs[-predict_last_n:], dim=1)
    return out @ w["linear.weight"].t() + w["linear.bias"]


def map_parameters(z, ranges: "OrderedDict[str, Tuple[float, float]]", dynamic: Sequence[str],
                   n_models: int, warmup: int):
    """z [B,T,P*m] (parameter index slow, member index fast) -> (warm-up dict, simulation dict)."""
    B, T, _ = z.shape
    z4 = z.reshape(B, T, -1, n_models)
    warm, sim = OrderedDict(), OrderedDict()
    for k, (name, (lo, hi)) in enumerate(ranges.items()):
        if name in dynamic:
            zw = z4[:, warmup - 1:warmup, k, :].expand(-1, warmup, -1)
            zs = z4[:, warmup:, k, :]
        else:
            zw = z4[:, -1:, k, :].expand(-1, warmup, -1)
            zs = z4[:, -1:, k, :].expand(-1, T - warmup, -1)
        warm[name] = lo + torch.sigmoid(zw) * (hi - lo)
        sim[name] = lo + torch.sigmoid(zs) * (hi - lo)
    return warm, sim


def _forcings(xc, m):
    p = xc[:, :, 0:1].expand(-1, -1, m)
    e = xc[:, :, 1:2].expand(-1, -1, m)
    if xc.shape[2] == 4:
        tmean = (xc[:, :, 2] + xc[:, :, 3]) / 2
    else:
        tmean = xc[:, :, 2]
    return p, e, tmean.unsqueeze(2).expand(-1, -1, m)


def _start_states(names, init, value, B, m, like):
    if init is not None:
        return [init[n] for n in names]
    return [like.new_full((B, m), value[n] if isinstance(value, dict) else value) for n in names]


# ----------------------------------------------------------------------------------------------
# SHM
# ----------------------------------------------------------------------------------------------
def shm_run(xc, par: Dict[str, torch.Tensor], n_models: int = 1, init: Optional[Dict] = None):
    B, T, _ = xc.shape
    m = n_models
    P, E, Tm = _forcings(xc, m)
    cold = Tm < 0
    zero = xc.new_zeros(())
    melt_pot = torch.where(cold, zero, Tm * par["dd"])
    rain = torch.where(cold, zero, P)
    snowfall = torch.where(cold, P, zero)
    pwp = 0.8 * par["sumax"]
    klu = 0.90
    ss, sf, su, si, sb = _start_states(SHM_STATES, init, SHM_INIT, B, m, xc)
    tape = {n: [] for n in SHM_STATES}
    q = []
    for j in range(T):
        fthr, smax, beta = par["f_thr"][:, j], par["sumax"][:, j], par["beta"][:, j]
        qs = torch.minimum(ss, melt_pot[:, j])
        ss = ss - qs + snowfall[:, j]
        qsp = qs + rain[:, j]
        qf_in = torch.maximum(zero, qsp - fthr)
        qu_in = torch.minimum(qsp, fthr)
        sf = sf + qf_in
        qf = sf * par["kf"][:, j]
        sf = sf - qf
        psi = (su / smax) ** beta
        su_t = su + qu_in * (1 - psi)
        su = torch.minimum(su_t, smax)
        qu = qu_in * psi + torch.maximum(zero, su_t - smax)
        kth = torch.where(su <= pwp[:, j], su / smax, torch.ones_like(su))
        su = torch.maximum(zero, su - E[:, j] * klu * kth)
        si = si + qu * par["perc"][:, j]
        qi = si * par["ki"][:, j]
        si = si - qi
        sb = sb + qu * (1.0 - par["perc"][:, j])
        qb = sb * par["kb"][:, j]
        sb = sb - qb
        for n, v in zip(SHM_STATES, (ss, sf, su, si, sb)):
            tape[n].append(v)
        q.append((qf + qi + qb).mean(dim=1))
    states = {n: torch.stack(v, dim=1) for n, v in tape.items()}
    return {
        "y_hat": torch.stack(q, dim=1).unsqueeze(2),
        "parameters": par,
        "internal_states": states,
        "final_states": {n: s[:, -1, :] for n, s in states.items()},
    }


# ----------------------------------------------------------------------------------------------
# HBV
# ----------------------------------------------------------------------------------------------
def hbv_run(xc, par: Dict[str, torch.Tensor], n_models: int = 1, init: Optional[Dict] = None):
    B, T, _ = xc.shape
    m = n_models
    P, E, Tm = _forcings(xc, m)
    zero = xc.new_zeros(())
    cold = Tm < par["TT"]
    rain = torch.where(cold, zero, P)
    snowfall = torch.where(cold, P, zero)
    SP, MW, SM, SUZ, SLZ = _start_states(HBV_STATES, init, HBV_INIT, B, m, xc)
    tape = {n: [] for n in HBV_STATES}
    q = []
    for j in range(T):
        TT, CFMAX, FC = par["TT"][:, j], par["CFMAX"][:, j], par["FC"][:, j]
        SP = SP + snowfall[:, j]
        melt = torch.clamp(CFMAX * (Tm[:, j] - TT), min=0.0)
        melt = torch.min(melt, SP)
        MW = MW + melt
        SP = SP - melt
        refr = torch.clamp(par["CFR"][:, j] * CFMAX * (TT - Tm[:, j]), min=0.0)
        refr = torch.min(refr, MW)
        SP = SP + refr
        MW = MW - refr
        tosoil = torch.clamp(MW - par["CWH"][:, j] * SP, min=0.0)
        MW = MW - tosoil
        wet = torch.clamp((SM / FC) ** par["BETA"][:, j], min=0.0, max=1.0)
        recharge = (rain[:, j] + tosoil) * wet
        SM = SM + rain[:, j] + tosoil - recharge
        excess = torch.clamp(SM - FC, min=0.0)
        SM = SM - excess
        if "BETAET" in par:
            ef = (SM / (par["LP"][:, j] * FC)) ** par["BETAET"][:, j]
        else:
            ef = SM / (par["LP"][:, j] * FC)
        ef = torch.clamp(ef, min=0.0, max=1.0)
        eta = torch.min(SM, E[:, j] * ef)
        SM = torch.clamp(SM - eta, min=1e-5)
        SUZ = SUZ + recharge + excess
        perc = torch.min(SUZ, par["PERC"][:, j])
        SUZ = SUZ - perc
        Q0 = par["K0"][:, j] * torch.clamp(SUZ - par["UZL"][:, j], min=0.0)
        SUZ = SUZ - Q0
        Q1 = par["K1"][:, j] * SUZ
        SUZ = SUZ - Q1
        SLZ = SLZ + perc
        Q2 = par["K2"][:, j] * SLZ
        SLZ = SLZ - Q2
        for n, v in zip(HBV_STATES, (SP, MW, SM, SUZ, SLZ)):
            tape[n].append(v)
        q.append((Q0 + Q1 + Q2).mean(dim=1))
    states = {n: torch.stack(v, dim=1) for n, v in tape.items()}
    return {
        "y_hat": torch.stack(q, dim=1).unsqueeze(2),
        "parameters": par,
        "internal_states": states,
        "final_states": {n: s[:, -1, :] for n, s in states.items()},
    }


# ----------------------------------------------------------------------------------------------
# linear reservoir, NonSense
# ----------------------------------------------------------------------------------------------
def _pack_result(q, tape, par):
    states = {n: torch.stack(v, dim=1) for n, v in tape.items()}
    return {
        "y_hat": torch.stack(q, dim=1).unsqueeze(2),
        "parameters": par,
        "internal_states": states,
        "final_states": {n: s[:, -1, :] for n, s in states.items()},
    }


def linres_run(xc, par: Dict[str, torch.Tensor], n_models: int = 1, init: Optional[Dict] = None):
    B, T, _ = xc.shape
    m = n_models
    zero = xc.new_zeros(())
    (si,) = _start_states(LINRES_STATES, init, LINRES_INIT, B, m, xc)
    tape, q = {"si": []}, []
    for j in range(T):
        p = xc[:, j, 0:1].expand(-1, m)
        et = xc[:, j, 1:2].expand(-1, m)
        si = torch.maximum(zero, (si + p) - et * par["aux_ET"][:, j])
        qi = si * par["ki"][:, j]
        si = si - qi
        tape["si"].append(si)
        q.append(qi.mean(dim=1))
    return _pack_result(q, tape, par)


def nonsense_run(xc, par: Dict[str, torch.Tensor], n_models: int = 1, init: Optional[Dict] = None):
    """n_models must be 1: the reference stores the [B, m] outflow in a [B, 1] slot (nonsense.py:163)."""
    B, T, _ = xc.shape
    m = n_models
    assert m == 1, "NonSense: the reference supports a single member"
    P = xc[:, :, 0:1].expand(-1, -1, m)
    E = xc[:, :, 1:2].expand(-1, -1, m)
    Tm = xc[:, :, 2:3].expand(-1, -1, m)          # column 2 as is, also with four columns (nonsense.py:80)
    cold = Tm < 0
    zero = xc.new_zeros(())
    melt_pot = torch.where(cold, zero, Tm * par["dd"])
    rain = torch.where(cold, zero, P)
    snowfall = torch.where(cold, P, zero)
    pwp = 0.8 * par["sumax"]
    klu = 0.90
    ss, su, si, sb = _start_states(NONSENSE_STATES, init, NONSENSE_INIT, B, m, xc)
    tape = {n: [] for n in NONSENSE_STATES}
    q = []
    for j in range(T):
        smax, beta = par["sumax"][:, j], par["beta"][:, j]
        qs = torch.minimum(ss, melt_pot[:, j])
        ss = ss - qs + snowfall[:, j]
        qsp = qs + rain[:, j]
        sb = sb + qsp
        qb = sb / par["kb"][:, j]
        sb = sb - qb
        si = si + qb
        qi = si / par["ki"][:, j]
        si = si - qi
        psi = (su / smax) ** beta
        su_t = su + qi * (1 - psi)
        su = torch.minimum(su_t, smax)
        qu = qi * psi + torch.maximum(zero, su_t - smax)
        kth = torch.where(su <= pwp[:, j], su / smax, torch.ones_like(su))
        su = torch.maximum(zero, su - E[:, j] * klu * kth)
        for n, v in zip(NONSENSE_STATES, (ss, su, si, sb)):
            tape[n].append(v)
        q.append(qu.mean(dim=1))
    return _pack_result(q, tape, par)


RUNNERS = {"shm": shm_run, "hbv": hbv_run, "linear_reservoir": linres_run, "nonsense": nonsense_run}


# ----------------------------------------------------------------------------------------------
# unit-hydrograph routing
# ----------------------------------------------------------------------------------------------
def uh_weights(alpha, beta, uh_len: int = UH_LEN):
    """Normalised gamma pdf sampled at k + 0.5, k = 0..uh_len-1. alpha, beta: [B] -> [uh_len, B]."""
    x = torch.arange(0.5, 0.5 + uh_len, 1, dtype=alpha.dtype, device=alpha.device).unsqueeze(1)
    coeff = 1 / (beta ** alpha * torch.lgamma(alpha).exp())
    pdf = coeff * (x ** (alpha - 1)) * torch.exp(-x / beta)
    return pdf / pdf.sum(dim=0)


def uh_route(q, alpha, beta, uh_len: int = UH_LEN):
    """q [B,T,1]; causal FIR y[t] = sum_k uh[k] q[t-k] with zero history."""
    uh = uh_weights(alpha, beta, uh_len)  # [K,B]
    B, T, _ = q.shape
    qpad = torch.nn.functional.pad(q[:, :, 0], (uh_len - 1, 0))
    y = q.new_zeros(B, T)
    for k in range(uh_len):
        y = y + uh[k].unsqueeze(1) * qpad[:, uh_len - 1 - k: uh_len - 1 - k + T]
    return y.unsqueeze(2)


# ----------------------------------------------------------------------------------------------
# hybrid chain
# ----------------------------------------------------------------------------------------------
def hybrid_forward(w: Dict[str, torch.Tensor], x_lstm, x_conc, *, model: str, n_models: int,
                   dynamic: Sequence[str], warmup: int, routing: bool = True, library: bool = False):
    ranges, _, _ = MODEL_TABLE[model]
    run = RUNNERS[model]
    f = lstm_library if library else lstm_cell_sequence
    h = f(x_lstm, w["lstm.weight_ih_l0"], w["lstm.weight_hh_l0"], w["lstm.bias_ih_l0"], w["lstm.bias_hh_l0"])
    z = h @ w["linear.weight"].t() + w["linear.bias"]
    npar = len(ranges) * n_models
    p_warm, p_sim = map_parameters(z[:, :, :npar], ranges, dynamic, n_models, warmup)
    with torch.no_grad():
        spin = run(x_conc[:, :warmup], p_warm, n_models)
    pred = run(x_conc[:, warmup:], p_sim, n_models, init=spin["final_states"])
    if routing:
        _, r = map_parameters(z[:, :, npar:], UH_RANGES, (), 1, warmup)
        pred["y_hat"] = uh_route(pred["y_hat"], r["alpha"][:, 0, 0], r["beta"][:, 0, 0])
    return pred


# ----------------------------------------------------------------------------------------------
# losses
# ----------------------------------------------------------------------------------------------
def nse_basin_averaged(y_sim, y_obs, per_basin_target_std):
    keep = ~torch.isnan(y_obs.flatten())
    err2 = (y_sim.flatten()[keep] - y_obs.flatten()[keep]) ** 2
    wgt = 1 / (per_basin_target_std.flatten()[keep] + 0.1) ** 2
    return torch.mean(wgt * err2)


def weighted_rmse(y_sim, y_obs):
    keep = ~torch.isnan(y_obs.flatten())
    s, o = y_sim.flatten()[keep], y_obs.flatten()[keep]
    ls = torch.log10(torch.sqrt(s + 1e-6) + 0.1)
    lo = torch.log10(torch.sqrt(o + 1e-6) + 0.1)
    return 0.75 * torch.sqrt(torch.mean((s - o) ** 2)) + 0.25 * torch.sqrt(torch.mean((ls - lo) ** 2))


# ----------------------------------------------------------------------------------------------
# optimiser tail: clip_grad_norm_(.., 1) then Adam (torch defaults), as the notebooks do
# ----------------------------------------------------------------------------------------------
def clip_and_adam(params: List[torch.Tensor], grads: List[torch.Tensor], exp_avg, exp_avg_sq, step: int,
                  lr: float = 1e-3, max_norm: float = 1.0, b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8):
    """In-place on params/exp_avg/exp_avg_sq; returns the pre-clip global norm."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads)).to(grads[0].dtype)
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    for p, g, m, v in zip(params, grads, exp_avg, exp_avg_sq):
        g = g * coef
        m.mul_(b1).add_(g, alpha=1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        p.addcdiv_(m, v.sqrt() / math.sqrt(bc2) + eps, value=-lr / bc1)
    return total


# ----------------------------------------------------------------------------------------------
# sampler
# ----------------------------------------------------------------------------------------------
def valid_window_flags(x, y, seq_length: int, predict_last_n: int):
    """x [N,nx], y [N,ny] numpy/torch with NaNs -> bool [N]: end index i gives a usable window."""
    x = torch.as_tensor(x)
    y = torch.as_tensor(y)
    N = x.shape[0]
    ok = torch.zeros(N, dtype=torch.bool)
    xn = torch.isnan(x).any(dim=1)
    yn = torch.isnan(y).all(dim=1)
    for i in range(seq_length - 1, N):
        if xn[i - seq_length + 1: i + 1].any():
            continue
        if yn[i - predict_last_n + 1: i + 1].all():
            continue
        ok[i] = True
    return ok


def window_gather(x_d, x_s, y_obs, x_conc, basin_std, row0, basin_idx, end_idx, seq_length, predict_last_n):
    """Resident concatenated series -> batch tensors, one (basin, end index) pair per sample.

    x_d [R,nd], y_obs [R,1], x_conc [R,nc] or None; x_s [nb,ns] or None; basin_std [nb]; row0 [nb]
    (first row of each basin in the concatenation).
    """
    xl, yo, xc, bs = [], [], [], []
    for b, i in zip(basin_idx.tolist(), end_idx.tolist()):
        r = int(row0[b]) + i
        win = x_d[r - seq_length + 1: r + 1]
        if x_s is not None:
            win = torch.cat([win, x_s[b].unsqueeze(0).expand(seq_length, -1)], dim=1)
        xl.append(win)
        yo.append(y_obs[r - predict_last_n + 1: r + 1])
        if x_conc is not None:
            xc.append(x_conc[r - seq_length + 1: r + 1])
        bs.append(basin_std[b].reshape(1, 1).expand(predict_last_n, 1))
    out = {"x_lstm": torch.stack(xl), "y_obs": torch.stack(yo), "basin_std": torch.stack(bs)}
    if xc:
        out["x_conceptual"] = torch.stack(xc)
    return out
