"""Single-step error of the gate pre-activations of the forward kernels: z recovered from the stored gates
(fp64 logit / atanh) against fp64 W_ih x_1 + W_hh h_0(kernel) + b.  Shows the sign and size of the tensor-core bias."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from hy2dl_b200 import _lib, ops
from hy2dl_b200.synthetic import lstm_weights
dev = torch.device("cuda:0")
# usage: python tools/dbg_zbias.py [H] [I]   (H = 128 / 256: the fp32 API path = streamed-weights tensor-core forward)
H = int(sys.argv[1]) if len(sys.argv) > 1 else 64
I = int(sys.argv[2]) if len(sys.argv) > 2 else 25
B, T = 4096, 2
rng = np.random.default_rng(5)
x = rng.normal(0, 1, (B, T, I)).astype(np.float32)
w = lstm_weights(I, H, 1, seed=4)
cu = lambda a: torch.tensor(a, device=dev)
wd = [cu(w[k]) for k in ("lstm.weight_ih_l0", "lstm.weight_hh_l0", "lstm.bias_ih_l0", "lstm.bias_hh_l0")]
G = (B + 31) // 32
W_ih, W_hh = w["lstm.weight_ih_l0"].astype(np.float64), w["lstm.weight_hh_l0"].astype(np.float64)
b = (w["lstm.bias_ih_l0"].astype(np.float64) + w["lstm.bias_hh_l0"].astype(np.float64))
for prec, name in (((_lib.PREC_FP32, "fp32"), (_lib.PREC_TF32X3, "tf32x3")) if H == 64 else ((_lib.PREC_FP32, "fp32 API, H=%d" % H),)):
    ri = prec == _lib.PREC_TF32X3
    x_in = ops.pack_x(cu(x), "ri32" if ri else "rowmajor")
    IPx = 32 if ri else (I + 7) // 8 * 8
    Bp = G * 32 if (ri or H >= 128) else B       # RI32 / RB8 tapes hold whole groups of 32 sequences
    h = torch.zeros(T, Bp if ri else B, H, device=dev); c = torch.zeros(T, Bp, H, device=dev); g = torch.zeros(T, Bp, 4 * H, device=dev)
    hl = torch.zeros(B, H, device=dev)
    ws = ops._ws(_lib.WS_LSTM_FWD, B, T, I, H, dev, prec)
    _lib.call("hy2dl_lstm_fwd", x_in.data_ptr(), B, T, I, IPx, H, *[t.data_ptr() for t in wd], h.data_ptr(), c.data_ptr(),
              g.data_ptr(), hl.data_ptr(), ws.data_ptr(), ws.numel(), prec, None, _lib.stream())
    torch.cuda.synchronize()
    if ri:
        h = ops.from_ri32(h.view(T, G, H // 32, 32, 32), B)
        g = ops.from_ri32(g.view(T, G, 4 * H // 32, 32, 32), B).reshape(T, B, H, 4).permute(0, 1, 3, 2).reshape(T, B, 4 * H)
    if not ri and H >= 128:      # RB8 [T][G][4H/8][32][8] -> row-major [T][B][4H]
        g = g.view(T, G, 4 * H // 8, 32, 8).permute(0, 1, 3, 2, 4).reshape(T, G * 32, 4 * H)[:, :B]
    h0 = h[0].cpu().double().numpy(); g1 = g[1].cpu().double().numpy()
    z_ref = x[:, 1].astype(np.float64) @ W_ih.T + h0 @ W_hh.T + b
    z_k = np.empty_like(z_ref)
    for gi in range(4):
        a = g1[:, gi * H:(gi + 1) * H]
        z_k[:, gi * H:(gi + 1) * H] = np.arctanh(a) if gi == 2 else np.log(a / (1 - a))
    m = (np.abs(z_ref) > 0.5) & (np.abs(z_ref) < 4)       # where the inverse activation is well conditioned
    rel = (z_k - z_ref) / np.abs(z_ref) * np.sign(z_ref)  # > 0: magnitude too large, < 0: too small
    print(f"{name}: signed relative error of |z|: mean {rel[m].mean():+.3e}  std {rel[m].std():.3e}   (n = {m.sum()})")
