"""2x2 experiment: forward tape from {fp32, tf32x3} kernel x backward by {fp32, tf32x3} kernel; weight gradients
contracted in fp64 on the host and compared with the fp64 oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from _util import O, tt
from hy2dl_b200 import _lib, ops
from hy2dl_b200.synthetic import lstm_weights
dev = torch.device("cuda:0")
B, T, I, H, dh_t0 = 96, 365, 25, 64, 183
rng = np.random.default_rng(3)
x = rng.normal(0, 1, (B, T, I)).astype(np.float32)
gh = rng.normal(0, 1, (T, B, H)).astype(np.float32); gh[:dh_t0] = 0
w = lstm_weights(I, H, 1, seed=4)
cu = lambda a: torch.tensor(a, device=dev)
wd = [cu(w[k]) for k in ("lstm.weight_ih_l0", "lstm.weight_hh_l0", "lstm.bias_ih_l0", "lstm.bias_hh_l0")]
G = (B + 31) // 32

def fwd(prec):
    ri = prec == _lib.PREC_TF32X3
    IP = 32 if ri else 32
    x_in = ops.pack_x(cu(x), "ri32" if ri else "rowmajor")
    IPx = 32 if ri else (I + 7) // 8 * 8
    Bp = G * 32 if ri else B
    h = torch.zeros(T, Bp, H, device=dev); c = torch.zeros(T, Bp, H, device=dev); g = torch.zeros(T, Bp, 4 * H, device=dev)
    hl = torch.zeros(B, H, device=dev)
    ws = ops._ws(_lib.WS_LSTM_FWD, B, T, I, H, dev, prec)
    _lib.call("hy2dl_lstm_fwd", x_in.data_ptr(), B, T, I, IPx, H, *[t.data_ptr() for t in wd], h.data_ptr(), c.data_ptr(),
              g.data_ptr(), hl.data_ptr(), ws.data_ptr(), ws.numel(), prec, None, _lib.stream())
    torch.cuda.synchronize()
    if ri:   # -> row-major, gate order g*H+u
        h = ops.from_ri32(h.view(T, G, H // 32, 32, 32), B); c = ops.from_ri32(c.view(T, G, H // 32, 32, 32), B)
        g = ops.from_ri32(g.view(T, G, 4 * H // 32, 32, 32), B).reshape(T, B, H, 4).permute(0, 1, 3, 2).reshape(T, B, 4 * H)
    return h.contiguous(), c.contiguous(), g.contiguous()

def bwd(prec, h, c, g):
    ri = prec == _lib.PREC_TF32X3
    dh = cu(gh)
    if ri:
        g_in = ops.to_ri32(g.reshape(T, B, 4, H).permute(0, 1, 3, 2).reshape(T, B, 4 * H).contiguous())
        c_in = ops.to_ri32(c); dh_in = ops.to_ri32(dh)
    else:
        g_in, c_in, dh_in = g.clone(), c, dh
    wsb = ops._ws(_lib.WS_LSTM_BWD, B, T, I, H, dev, prec)
    _lib.call("hy2dl_lstm_bwd", B, T, H, wd[1].data_ptr(), g_in.data_ptr(), c_in.data_ptr(), dh_in.data_ptr(), dh_t0, None,
              g_in.data_ptr(), wsb.data_ptr(), wsb.numel(), prec, None, _lib.stream())
    torch.cuda.synchronize()
    if ri:
        dG = ops.from_ri32(g_in.view(T, G, 4 * H // 32, 32, 32), B).reshape(T, B, H, 4).permute(0, 1, 3, 2).reshape(T, B, 4 * H)
    else:
        dG = g_in
    return dG.cpu().double().numpy()

wd64 = {k: tt(v, torch.float64, True) for k, v in w.items() if k.startswith("lstm.")}
ref = O.lstm_cell_sequence(tt(x, torch.float64), wd64["lstm.weight_ih_l0"], wd64["lstm.weight_hh_l0"], wd64["lstm.bias_ih_l0"], wd64["lstm.bias_hh_l0"])
(ref * tt(gh.transpose(1, 0, 2), torch.float64)).sum().backward()
R = {k: v.grad.numpy() for k, v in wd64.items()}
names = {_lib.PREC_FP32: "fp32", _lib.PREC_TF32X3: "tf32x3"}
for pf in (_lib.PREC_FP32, _lib.PREC_TF32X3):
    h, c, g = fwd(pf)
    herr = np.abs(h.cpu().double().numpy().transpose(1, 0, 2) - ref.detach().numpy()).max()
    for pb in (_lib.PREC_FP32, _lib.PREC_TF32X3):
        dG = bwd(pb, h, c, g)
        hh = h.cpu().double().numpy()
        hprev = np.concatenate([np.zeros_like(hh[:1]), hh[:-1]], axis=0)
        dw_hh = np.einsum("tbm,tbk->mk", dG, hprev); dw_ih = np.einsum("tbm,bti->mi", dG, x.astype(np.float64)); db = dG.sum((0, 1))
        e = [np.abs(a - R[k]).max() / np.abs(R[k]).max() for a, k in ((dw_ih, "lstm.weight_ih_l0"), (dw_hh, "lstm.weight_hh_l0"), (db, "lstm.bias_ih_l0"))]
        print(f"fwd {names[pf]:7s} (h err {herr:.2e})  bwd {names[pb]:7s}: dW_ih {e[0]:.2e} dW_hh {e[1]:.2e} db {e[2]:.2e}")
