"""Tensor-core (tcgen05, 3xTF32) LSTM path against the fp64 oracle, through the C ABI.
Tolerance: the same rtol 1e-4 as the fp32 path (the split keeps ~21 mantissa bits per product)."""
import numpy as np
import pytest
import torch

from _util import O, assert_close, tt

pytestmark = pytest.mark.gpu


def _run_fwd(dev, B, T, I, precision, seed=0, tape=True):
    from hy2dl_b200 import _lib, ops
    from hy2dl_b200.synthetic import lstm_weights
    H, IP = 64, (I + 7) // 8 * 8
    rng = np.random.default_rng(seed)
    x = rng.normal(0, 1, (B, T, I)).astype(np.float32)
    w = lstm_weights(I, H, 1, seed=seed + 1)
    cu = lambda a: torch.tensor(a, device=dev)
    xT = ops.pack_x(cu(x))                               # [T,B,IP] row-major
    x_in = ops.to_ri32(xT) if precision == _lib.PREC_TF32X3 else xT
    G = (B + 31) // 32
    Bp = G * 32 if precision == _lib.PREC_TF32X3 else B
    h = torch.zeros(T, Bp, H, device=dev)
    c = torch.zeros(T, Bp, H, device=dev) if tape else None
    g = torch.zeros(T, Bp, 4 * H, device=dev) if tape else None
    hl = torch.zeros(B, H, device=dev)
    ws = ops._ws(_lib.WS_LSTM_FWD, B, T, I, H, dev, precision)
    wd = [cu(w[k]) for k in ("lstm.weight_ih_l0", "lstm.weight_hh_l0", "lstm.bias_ih_l0", "lstm.bias_hh_l0")]
    _lib.call("hy2dl_lstm_fwd", x_in.data_ptr(), B, T, I, IP, H, *[t.data_ptr() for t in wd], h.data_ptr(),
              None if c is None else c.data_ptr(), None if g is None else g.data_ptr(), hl.data_ptr(), ws.data_ptr(),
              ws.numel(), precision, _lib.stream())
    torch.cuda.synchronize()
    if precision == _lib.PREC_TF32X3:
        unri = lambda a, F: ops.from_ri32(a.view(T, G, F // 4, 32, 4), B)
        h = unri(h, H)
        if tape:
            c = unri(c, H)
            g = unri(g, 4 * H).reshape(T, B, H, 4).permute(0, 1, 3, 2).reshape(T, B, 4 * H)   # n = 4u+g -> g*H+u
    wd64 = {k: tt(v, torch.float64) for k, v in w.items()}
    ref = O.lstm_cell_sequence(tt(x, torch.float64), wd64["lstm.weight_ih_l0"], wd64["lstm.weight_hh_l0"],
                               wd64["lstm.bias_ih_l0"], wd64["lstm.bias_hh_l0"]).numpy()   # [B,T,H]
    return h.cpu().numpy(), (None if c is None else c.cpu().numpy()), (None if g is None else g.cpu().numpy()), \
        hl.cpu().numpy(), ref, x, w


@pytest.mark.parametrize("B,T,I", [(128, 4, 25), (200, 12, 25), (64, 365, 25), (300, 30, 8)])
def test_tc_fwd_vs_oracle(B, T, I):
    from hy2dl_b200 import _lib
    dev = torch.device("cuda:0")
    h, c, g, hl, ref, x, w = _run_fwd(dev, B, T, I, _lib.PREC_TF32X3)
    hf, cf, gf, _, _, _, _ = _run_fwd(dev, B, T, I, _lib.PREC_FP32)
    for t in range(min(T, 4)):
        print(f"t={t}: tc err {np.abs(h[t] - ref[:, t]).max():.3e}  fp32-kernel err {np.abs(hf[t] - ref[:, t]).max():.3e}  "
              f"gate err vs fp32 kernel {np.abs(g[t] - gf[t]).max():.3e}")
    assert_close(g[0], gf[0], rtol=1e-5, atol=1e-6, what="gates t=0 (tc vs fp32 kernel)")
    assert_close(h.transpose(1, 0, 2), ref, rtol=1e-4, atol=1e-6, what="h_seq")
    assert_close(hl, ref[:, -1], rtol=1e-4, atol=1e-6, what="h_last")
    assert_close(c, cf, rtol=1e-4, atol=1e-6, what="c_seq vs fp32 kernel")
