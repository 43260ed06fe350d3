"""Which kernel of the H = 128 / 256 chain carries the gradient error?  Runs fwd / BPTT / wgrad through the C ABI on a
golden CudaLSTM case, once on the chain's own tapes and once on exact (fp64 oracle, rounded to fp32) tapes, and compares
every intermediate with the fp64 oracle.   python tools/dbg_chain_isolate.py [case]"""
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
P = _lib.PREC_FP32

# ---- fp64 oracle with intermediates
d = lambda a: torch.tensor(np.asarray(a), dtype=torch.float64)
x = d(g["x_lstm"])
w_ih, w_hh = d(w["lstm.weight_ih_l0"]).requires_grad_(), d(w["lstm.weight_hh_l0"]).requires_grad_()
b_ih, b_hh = d(w["lstm.bias_ih_l0"]).requires_grad_(), d(w["lstm.bias_hh_l0"]).requires_grad_()
lw, lb = d(w["linear.weight"]), d(w["linear.bias"])
h = torch.zeros(B, H, dtype=torch.float64); c = torch.zeros(B, H, dtype=torch.float64)
zs, hs, cs, gs = [], [], [], []
for t in range(T):
    z = x[:, t] @ w_ih.t() + b_ih + h @ w_hh.t() + b_hh
    z.retain_grad(); zs.append(z)
    i, f, gg, o = z[:, :H].sigmoid(), z[:, H:2 * H].sigmoid(), z[:, 2 * H:3 * H].tanh(), z[:, 3 * H:].sigmoid()
    c = f * c + i * gg
    h = o * c.tanh()
    hs.append(h); cs.append(c); gs.append(torch.cat([i, f, gg, o], 1))
y = h @ lw.t() + lb
yo, sd = d(g["y_obs"])[:, 0], d(g["basin_std"])[:, 0]
m = ~torch.isnan(yo)
loss = (((y - torch.nan_to_num(yo)) ** 2 / (sd + 0.1) ** 2)[m]).mean()
h.retain_grad()
loss.backward()
dh_last_ref = h.grad
H_ref = torch.stack([v.detach() for v in hs]); C_ref = torch.stack([v.detach() for v in cs]); G_ref = torch.stack([v.detach() for v in gs])
dZ_ref = torch.stack([z.grad for z in zs])
rel = lambda a, b: float((a.double().cpu() - b).abs().max() / b.abs().max())

def to_rb8(a):   # [T,B,F] -> RB8 [T,G,F/8,32,8]
    Tn, Bn, F = a.shape
    G = (Bn + 31) // 32
    p = torch.zeros(Tn, G * 32, F, dtype=a.dtype, device=a.device); p[:, :Bn] = a
    return p.view(Tn, G, 32, F // 8, 8).permute(0, 1, 3, 2, 4).contiguous()

def from_rb8(a, Bn):
    Tn, G, FB = a.shape[:3]
    return a.permute(0, 1, 3, 2, 4).reshape(Tn, G * 32, FB * 8)[:, :Bn]

f32 = lambda a: torch.tensor(np.asarray(a), dtype=torch.float32, device=dev)
xT = ops.pack_x(f32(g["x_lstm"]))
IP = xT.shape[2]
W = [f32(w[k]) for k in ("lstm.weight_ih_l0", "lstm.weight_hh_l0", "lstm.bias_ih_l0", "lstm.bias_hh_l0")]
G32 = (B + 31) // 32
h_seq = torch.empty(T, B, H, device=dev); h_last = torch.empty(B, H, device=dev)
c_seq = torch.zeros(T, G32, H // 8, 32, 8, device=dev); gates = torch.zeros(T, G32, 4 * H // 8, 32, 8, device=dev)
ws = ops._ws(_lib.WS_LSTM_FWD, B, T, I, H, dev, P)
call("hy2dl_lstm_fwd", ptr(xT), B, T, I, IP, H, *[ptr(t) for t in W], ptr(h_seq), ptr(c_seq), ptr(gates), ptr(h_last), ptr(ws), ws.numel(), P, None, stream())
print("fwd: h", rel(h_seq, H_ref), "c", rel(from_rb8(c_seq, B), C_ref), "gates", rel(from_rb8(gates, B), G_ref))

def bwd_wgrad(tag, gates_t, c_t, h_t):
    dG = gates_t.clone()
    dhl = dh_last_ref.float().to(dev).contiguous()
    wsb = ops._ws(_lib.WS_LSTM_BWD, B, T, I, H, dev, P)
    call("hy2dl_lstm_bwd", B, T, H, ptr(W[1]), ptr(dG), ptr(c_t), None, 0, ptr(dhl), ptr(dG), ptr(wsb), wsb.numel(), P, None, stream())
    dz = from_rb8(dG, B)
    per_t = (dz.double().cpu() - dZ_ref).abs().amax((1, 2)) / dZ_ref.abs().amax((1, 2))
    print(tag, "BPTT dG rel err (global)", rel(dz, dZ_ref), "per-step rel err at t = T-1, T-10, T-50, T-100, T-200, 0:",
          [float(per_t[t]) for t in (T - 1, T - 10, T - 50, T - 100, T - 200, 0)])
    for src_name, dG_src in (("own dG", dG), ("exact dG", to_rb8(dZ_ref.float().to(dev)))):
        dw_ih = torch.empty(4 * H, I, device=dev); dw_hh = torch.empty(4 * H, H, device=dev); db = torch.empty(4 * H, device=dev)
        wsw = ops._ws(_lib.WS_LSTM_WGRAD, B, T, I, H, dev, P)
        call("hy2dl_lstm_wgrad", ptr(dG_src), ptr(h_t), ptr(xT), B, T, I, IP, H, ptr(dw_ih), ptr(dw_hh), ptr(db), ptr(wsw), wsw.numel(), P, None, stream())
        print(tag, "wgrad on", src_name, ": dW_ih", rel(dw_ih, w_ih.grad), "dW_hh", rel(dw_hh, w_hh.grad), "db", rel(db, b_ih.grad))

bwd_wgrad("[own tapes]", gates, c_seq, h_seq)
bwd_wgrad("[exact tapes]", to_rb8(G_ref.float().to(dev)), to_rb8(C_ref.float().to(dev)), H_ref.float().to(dev).contiguous())

# ---- forward error anatomy: per gate and per time step (abs error of the post-activation gate values, c and h)
Gm = from_rb8(gates, B).double().cpu(); Cm = from_rb8(c_seq, B).double().cpu(); Hm = h_seq.double().cpu()
for tt_ in (0, 1, 2, 10, 100, T - 1):
    e = Gm[tt_] - G_ref[tt_]
    per = [(float(e[:, k * H:(k + 1) * H].abs().max()), float(e[:, k * H:(k + 1) * H].mean())) for k in range(4)]
    print(f"t={tt_}: gate abs err max/mean i,f,g,o:", ["%.1e/%+.1e" % p for p in per],
          "c max %.1e mean %+.1e (|c| max %.2f)" % (float((Cm[tt_] - C_ref[tt_]).abs().max()), float((Cm[tt_] - C_ref[tt_]).mean()), float(C_ref[tt_].abs().max())),
          "h max %.1e" % float((Hm[tt_] - H_ref[tt_]).abs().max()))
# implied pre-activation error of the sigmoid gates: dz = dgate / (gate (1 - gate))
for k, nm in ((0, "i"), (1, "f"), (3, "o")):
    gv = G_ref[:, :, k * H:(k + 1) * H]
    dz = (Gm[:, :, k * H:(k + 1) * H] - gv) / (gv * (1 - gv)).clamp_min(1e-6)
    zt = torch.stack([z.detach() for z in zs])[:, :, k * H:(k + 1) * H]
    relz = dz / zt.abs().clamp_min(0.5)
    print(f"gate {nm}: implied dz max {float(dz.abs().max()):.2e} mean {float(dz.mean()):+.2e} std {float(dz.std()):.2e}; "
          f"dz/|z| mean {float(relz.mean()):+.2e} std {float(relz.std()):.2e}; corr(dz, z) {float(torch.corrcoef(torch.stack([dz.flatten(), zt.flatten()]))[0, 1]):+.3f}")
