"""Intrinsic per-step error of the recurrent forward kernels: every step of OUR tape is re-evaluated in fp64 from OUR
h_{t-1}, c_{t-1} (teacher forcing), so the recurrence's error amplification is excluded and what remains is the
arithmetic error one step adds (pre-activation z, cell state, hidden state).  HY2DL_FP32_TC=0: the CUDA-core chain.
    python tools/dbg_step_noise.py [case]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _util import cudalstm_weights_np, golden  # noqa: E402

from hy2dl_b200 import _lib, ops  # noqa: E402
from hy2dl_b200._lib import call, ptr, stream  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cudalstm_c3_full"
g = golden(name)
B, T, I, H, wseed = [int(v) for v in g["meta"]][:5]
dev = torch.device("cuda:0")
w = cudalstm_weights_np(g)
tc = os.environ.get("HY2DL_FP32_TC", "1") != "0"
prec = sys.argv[2] if len(sys.argv) > 2 else "fp32"
P = ops.PRECISIONS[prec]
ri = prec != "fp32" and H == 64
f32 = lambda a: torch.tensor(np.asarray(a), dtype=torch.float32, device=dev)
xT = ops.pack_x(f32(g["x_lstm"]), "ri32" if ri else "rowmajor")
IP = 32 if ri else xT.shape[2]
W = [f32(w[k]) for k in ("lstm.weight_ih_l0", "lstm.weight_hh_l0", "lstm.bias_ih_l0", "lstm.bias_hh_l0")]
G32 = (B + 31) // 32


def from_rb8(a, Bn):
    Tn, G, FB = a.shape[:3]
    return a.permute(0, 1, 3, 2, 4).reshape(Tn, G * 32, FB * 8)[:, :Bn]


if ri:
    seq = lambda F: torch.zeros(T, G32, F // 32, 32, 32, device=dev)
    h_seq, c_seq, gates = seq(H), seq(H), seq(4 * H)
elif tc and H >= 128:
    h_seq = torch.empty(T, B, H, device=dev)
    c_seq = torch.zeros(T, G32, H // 8, 32, 8, device=dev); gates = torch.zeros(T, G32, 4 * H // 8, 32, 8, device=dev)
else:
    h_seq = torch.empty(T, B, H, device=dev); c_seq = torch.empty(T, B, H, device=dev); gates = torch.empty(T, B, 4 * H, device=dev)
h_last = torch.empty(B, H, device=dev)
ws = ops._ws(_lib.WS_LSTM_FWD, B, T, I, H, dev, P)
call("hy2dl_lstm_fwd", ptr(xT), B, T, I, IP, H, *[ptr(t) for t in W], ptr(h_seq), ptr(c_seq), ptr(gates), ptr(h_last), ptr(ws), ws.numel(), P, None, stream())
if ri:
    Hm, Cm = ops.from_ri32(h_seq, B), ops.from_ri32(c_seq, B)
    Gm = ops.from_ri32(gates, B).view(T, B, H, 4).permute(0, 1, 3, 2).reshape(T, B, 4 * H)     # n = 4u + g -> g*H + u
elif tc and H >= 128:
    Hm, Cm, Gm = h_seq, from_rb8(c_seq, B), from_rb8(gates, B)
else:
    Hm, Cm, Gm = h_seq, c_seq, gates
Hm, Cm, Gm = Hm.double().cpu(), Cm.double().cpu(), Gm.double().cpu()
d = lambda a: torch.tensor(np.asarray(a), dtype=torch.float64)
x = d(g["x_lstm"]); w_ih, w_hh = d(w["lstm.weight_ih_l0"]), d(w["lstm.weight_hh_l0"]); b = d(w["lstm.bias_ih_l0"]) + d(w["lstm.bias_hh_l0"])
hp = torch.cat([torch.zeros(1, B, H, dtype=torch.float64), Hm[:-1]]); cp = torch.cat([torch.zeros(1, B, H, dtype=torch.float64), Cm[:-1]])
z = torch.einsum("bti,ni->tbn", x, w_ih) + b + torch.einsum("tbh,nh->tbn", hp, w_hh)
gi, gf, gg, go = z[..., :H].sigmoid(), z[..., H:2 * H].sigmoid(), z[..., 2 * H:3 * H].tanh(), z[..., 3 * H:].sigmoid()
c_ref = gf * cp + gi * gg
h_ref = go * c_ref.tanh()
print({"case": name, "chain": prec if ri else ("tcgen05" if tc and H >= 128 else "cuda-core"), "H": H, "max|c|": float(Cm.abs().max())})
for k, nm, ref in ((0, "i", gi), (1, "f", gf), (2, "g", gg), (3, "o", go)):
    e = Gm[..., k * H:(k + 1) * H] - ref
    der = (ref * (1 - ref)) if k != 2 else (1 - ref * ref)
    dz = e / der.clamp_min(1e-3)
    zz = z[..., k * H:(k + 1) * H]
    print(f"  gate {nm}: abs err max {float(e.abs().max()):.2e} rms {float(e.pow(2).mean().sqrt()):.2e} | implied dz rms {float(dz.pow(2).mean().sqrt()):.2e} "
          f"mean {float(dz.mean()):+.2e} | dz/|z| (|z|>0.5) mean {float((dz / zz.abs())[zz.abs() > 0.5].mean()):+.2e} | |z| rms {float(zz.pow(2).mean().sqrt()):.2f}")
ec, eh = Cm - c_ref, Hm - h_ref
print(f"  c: one-step err max {float(ec.abs().max()):.2e} rms {float(ec.pow(2).mean().sqrt()):.2e};  h: max {float(eh.abs().max()):.2e} rms {float(eh.pow(2).mean().sqrt()):.2e}")
# persistence: rms over (b, u) of the time-mean error, against rms / sqrt(T) expected of step-to-step independent errors
def persist(e, nm):
    tm = e.mean(0)
    print(f"  persistence {nm}: rms of time-mean {float(tm.pow(2).mean().sqrt()):.2e} vs independent-noise expectation {float(e.pow(2).mean().sqrt()) / np.sqrt(e.shape[0]):.2e}; global mean {float(e.mean()):+.2e}")
persist(ec, "c one-step")
persist(eh, "h one-step")
for k, nm, ref in ((0, "i", gi), (1, "f", gf), (2, "g", gg), (3, "o", go)):
    persist(Gm[..., k * H:(k + 1) * H] - ref, "gate " + nm)
