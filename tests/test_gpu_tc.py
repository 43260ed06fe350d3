"""Tensor-core (tcgen05, 3xTF32) LSTM path against the fp64 oracle, through the C ABI.
Tolerance: the same rtol 1e-4 as the fp32 path (the split keeps ~21 mantissa bits per product)."""
import numpy as np
import pytest
import torch

from _util import O, assert_close, tt

pytestmark = pytest.mark.gpu


def _unrb8(a, B):
    """RB8 tape [T][ceil(B/32)][F/8][32 rows][8] (stored in a [T, Bp, F] buffer) -> row-major [T, B, F].  The fp32-class
    kernels keep their internal tapes (c, gates / dG) in this layout: H = 128 / 256 always, H = 64 at the small batches that
    run on the unit-split kernels (every batch of this file)."""
    T, Bp, F = a.shape
    return a.view(T, Bp // 32, F // 8, 32, 8).permute(0, 1, 3, 2, 4).reshape(T, Bp, F)[:, :B]


def _run_fwd(dev, B, T, I, precision, seed=0, tape=True):
    from hy2dl_b200 import _lib, ops
    from hy2dl_b200.synthetic import lstm_weights
    ri = precision == _lib.PREC_TF32X3
    H, IP = 64, 32 if ri else (I + 7) // 8 * 8
    rng = np.random.default_rng(seed)
    x = rng.normal(0, 1, (B, T, I)).astype(np.float32)
    w = lstm_weights(I, H, 1, seed=seed + 1)
    cu = lambda a: torch.tensor(a, device=dev)
    x_in = ops.pack_x(cu(x), "ri32" if ri else "rowmajor")
    G = (B + 31) // 32
    Bp = G * 32                                                     # tapes hold whole groups of 32 sequences
    h = torch.zeros(T, Bp if ri else B, H, device=dev)
    c = torch.zeros(T, Bp, H, device=dev) if tape else None
    g = torch.zeros(T, Bp, 4 * H, device=dev) if tape else None
    hl = torch.zeros(B, H, device=dev)
    ws = ops._ws(_lib.WS_LSTM_FWD, B, T, I, H, dev, precision)
    wd = [cu(w[k]) for k in ("lstm.weight_ih_l0", "lstm.weight_hh_l0", "lstm.bias_ih_l0", "lstm.bias_hh_l0")]
    _lib.call("hy2dl_lstm_fwd", x_in.data_ptr(), B, T, I, IP, H, *[t.data_ptr() for t in wd], h.data_ptr(),
              None if c is None else c.data_ptr(), None if g is None else g.data_ptr(), hl.data_ptr(), ws.data_ptr(),
              ws.numel(), precision, None, _lib.stream())
    torch.cuda.synchronize()
    if precision == _lib.PREC_TF32X3:
        unri = lambda a, F: ops.from_ri32(a.view(T, G, F // 32, 32, 32), B)
        h = unri(h, H)
        if tape:
            c = unri(c, H)
            g = unri(g, 4 * H).reshape(T, B, H, 4).permute(0, 1, 3, 2).reshape(T, B, 4 * H)   # n = 4u+g -> g*H+u
    elif tape:
        c, g = _unrb8(c, B), _unrb8(g, B)
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


def _bwd_both(dev, B, T, I, dh_t0, use_last, seed=3):
    """Run fwd+bwd with both precisions on the same inputs; return row-major dG (g*H+u order) of each."""
    from hy2dl_b200 import _lib, ops
    from hy2dl_b200.synthetic import lstm_weights
    H = 64
    rng = np.random.default_rng(seed)
    x = rng.normal(0, 1, (B, T, I)).astype(np.float32)
    gh = rng.normal(0, 1, (T, B, H)).astype(np.float32)
    gh[:dh_t0] = 0
    gl = rng.normal(0, 1, (B, H)).astype(np.float32)
    w = lstm_weights(I, H, 1, seed=seed + 1)
    cu = lambda a: torch.tensor(a, device=dev)
    wd = [cu(w[k]) for k in ("lstm.weight_ih_l0", "lstm.weight_hh_l0", "lstm.bias_ih_l0", "lstm.bias_hh_l0")]
    G = (B + 31) // 32
    out, wg = {}, {}
    for prec in (_lib.PREC_FP32, _lib.PREC_TF32X3):
        ri = prec == _lib.PREC_TF32X3
        Bp = G * 32
        IP = 32 if ri else (I + 7) // 8 * 8
        x_in = ops.pack_x(cu(x), "ri32" if ri else "rowmajor")
        h = torch.zeros(T, Bp if ri else B, H, device=dev); c = torch.zeros(T, Bp, H, device=dev)
        g = torch.zeros(T, Bp, 4 * H, device=dev); hl = torch.zeros(B, H, device=dev)
        ws = ops._ws(_lib.WS_LSTM_FWD, B, T, I, H, dev, prec)
        _lib.call("hy2dl_lstm_fwd", x_in.data_ptr(), B, T, I, IP, H, *[t.data_ptr() for t in wd], h.data_ptr(),
                  c.data_ptr(), g.data_ptr(), hl.data_ptr(), ws.data_ptr(), ws.numel(), prec, None, _lib.stream())
        dh = cu(gh)
        dh_in = ops.to_ri32(dh) if ri else dh
        dl = cu(gl) if use_last else None
        wsb = ops._ws(_lib.WS_LSTM_BWD, B, T, I, H, dev, prec)
        _lib.call("hy2dl_lstm_bwd", B, T, H, wd[1].data_ptr(), g.data_ptr(), c.data_ptr(), dh_in.data_ptr(), dh_t0,
                  None if dl is None else dl.data_ptr(), g.data_ptr(), wsb.data_ptr(), wsb.numel(), prec, None, _lib.stream())
        # weight gradients with the same precision's kernel
        dwi = torch.full((4 * H, I), np.nan, device=dev); dwh = torch.full((4 * H, H), np.nan, device=dev)
        dbb = torch.full((4 * H,), np.nan, device=dev)
        wsw = ops._ws(_lib.WS_LSTM_WGRAD, B, T, I, H, dev, prec)
        _lib.call("hy2dl_lstm_wgrad", g.data_ptr(), h.data_ptr(), x_in.data_ptr(), B, T, I, IP, H, dwi.data_ptr(),
                  dwh.data_ptr(), dbb.data_ptr(), wsw.data_ptr(), wsw.numel(), prec, None, _lib.stream())
        torch.cuda.synchronize()
        wg[prec] = {"lstm.weight_ih_l0": dwi.cpu().numpy(), "lstm.weight_hh_l0": dwh.cpu().numpy(),
                    "lstm.bias_ih_l0": dbb.cpu().numpy()}
        if ri:
            dG = ops.from_ri32(g.view(T, G, 4 * H // 32, 32, 32), B).reshape(T, B, H, 4).permute(0, 1, 3, 2).reshape(T, B, 4 * H)
            hh = ops.from_ri32(h.view(T, G, H // 32, 32, 32), B)
        else:
            dG, hh = _unrb8(g, B), h
        out[prec] = (dG.cpu().double().numpy(), hh.cpu().double().numpy())
    # oracle gradients (fp64 autograd)
    wd64 = {k: tt(v, torch.float64, True) for k, v in w.items() if k.startswith("lstm.")}
    ref = O.lstm_cell_sequence(tt(x, torch.float64), wd64["lstm.weight_ih_l0"], wd64["lstm.weight_hh_l0"],
                               wd64["lstm.bias_ih_l0"], wd64["lstm.bias_hh_l0"])
    f = (ref * tt(gh.transpose(1, 0, 2), torch.float64)).sum()
    if use_last:
        f = f + (ref[:, -1] * tt(gl, torch.float64)).sum()
    f.backward()
    return out, {k: v.grad.numpy() for k, v in wd64.items()}, x, wg


@pytest.mark.parametrize("B,T,I,dh_t0,use_last", [(128, 3, 25, 0, False), (200, 40, 25, 17, True), (96, 365, 25, 183, False)])
def test_tc_bwd_vs_fp32_kernel_and_oracle(B, T, I, dh_t0, use_last):
    from hy2dl_b200 import _lib
    dev = torch.device("cuda:0")
    out, ref, x, wg = _bwd_both(dev, B, T, I, dh_t0, use_last)
    dG32, h32 = out[_lib.PREC_FP32]
    dGtc, htc = out[_lib.PREC_TF32X3]
    for t in sorted({T - 1, max(T - 2, 0), 0}, reverse=True):
        print(f"t={t}: |dG_tc - dG_fp32| max {np.abs(dGtc[t] - dG32[t]).max():.3e} (scale {np.abs(dG32[t]).max():.3e})")
    # kernel against kernel: both carry their own rounding through T recurrent steps (and the backward
    # pass re-amplifies it), so the two tapes drift apart with T; the parity claims are the oracle checks below
    long_seq = T >= 100
    assert_close(dGtc[T - 1], dG32[T - 1], rtol=1e-4 if long_seq else 1e-5, atol=1e-7, what="dG at the last step")
    assert_close(dGtc, dG32, rtol=5e-4 if long_seq else 1e-4, atol=1e-7, what="dG all steps (tc vs fp32 kernel)")
    # weight gradients from the tc tape, contracted in fp64 on the host, against oracle autograd
    hprev = np.concatenate([np.zeros_like(htc[:1]), htc[:-1]], axis=0)
    dw_hh = np.einsum("tbm,tbk->mk", dGtc, hprev)
    db = dGtc.sum((0, 1))
    dw_ih = np.einsum("tbm,bti->mi", dGtc, x.astype(np.float64))
    assert_close(dw_hh, ref["lstm.weight_hh_l0"], rtol=1e-4, atol=1e-7, what="dW_hh")
    assert_close(dw_ih, ref["lstm.weight_ih_l0"], rtol=1e-4, atol=1e-7, what="dW_ih")
    assert_close(db, ref["lstm.bias_ih_l0"], rtol=1e-4, atol=1e-7, what="db")
    # the tcgen05 weight-gradient GEMM (RI32 slabs as MN-major UMMA operands) and the fp32 one
    for prec, name in ((_lib.PREC_FP32, "fp32"), (_lib.PREC_TF32X3, "tf32x3")):
        for k in ("lstm.weight_hh_l0", "lstm.weight_ih_l0", "lstm.bias_ih_l0"):
            assert_close(wg[prec][k], ref[k], rtol=1e-4, atol=1e-7, what=f"wgrad kernel ({name}) {k}")


@pytest.mark.parametrize("B,T,I", [(32, 1, 8), (70, 5, 31), (4800, 3, 16)])
def test_tc_wgrad_shapes(B, T, I):
    """Ragged groups, several input sizes, more groups than CTAs; checked against an fp64 contraction."""
    from hy2dl_b200 import _lib, ops
    dev = torch.device("cuda:0")
    H, IP, G = 64, 32, (B + 31) // 32
    rng = np.random.default_rng(B + I)
    dG = rng.normal(0, 1, (T, B, 4 * H)).astype(np.float32)       # columns n = 4u + g
    h = rng.normal(0, 1, (T, B, H)).astype(np.float32)
    x = np.zeros((T, B, IP), np.float32); x[:, :, :I] = rng.normal(0, 1, (T, B, I))
    cu = lambda a: torch.tensor(a, device=dev)
    dwi = torch.full((4 * H, I), np.nan, device=dev); dwh = torch.full((4 * H, H), np.nan, device=dev)
    dbb = torch.full((4 * H,), np.nan, device=dev)
    ws = ops._ws(_lib.WS_LSTM_WGRAD, B, T, I, H, dev, _lib.PREC_TF32X3)
    args = [ops.to_ri32(cu(a)) for a in (dG, h, x)]
    _lib.call("hy2dl_lstm_wgrad", *[a.data_ptr() for a in args], B, T, I, IP, H, dwi.data_ptr(), dwh.data_ptr(),
              dbb.data_ptr(), ws.data_ptr(), ws.numel(), _lib.PREC_TF32X3, None, _lib.stream())
    torch.cuda.synchronize()
    d = dG.astype(np.float64).reshape(T, B, H, 4).transpose(0, 1, 3, 2).reshape(T, B, 4 * H)   # -> g*H + u
    hprev = np.concatenate([np.zeros_like(h[:1]), h[:-1]]).astype(np.float64)
    assert_close(dwi.cpu().numpy(), np.einsum("tbm,tbi->mi", d, x[:, :, :I].astype(np.float64)), rtol=2e-6, what="dW_ih")
    assert_close(dwh.cpu().numpy(), np.einsum("tbm,tbk->mk", d, hprev), rtol=2e-6, atol=1e-30, what="dW_hh")
    assert_close(dbb.cpu().numpy(), d.sum((0, 1)), rtol=2e-6, what="db")


@pytest.mark.parametrize("B,T,I,H", [(3, 5, 8, 128), (40, 7, 64, 256), (33, 3, 32, 128), (64, 2, 16, 256), (700, 9, 41, 256),
                                     (1, 1, 31, 128)])
def test_rowmajor_tc_wgrad_shapes(B, T, I, H):
    """Weight gradients of the fp32 path for H = 128 / 256 (tcgen05 GEMM on the row-major tapes): input sizes whose
    ones column falls into the zero-filled padding (I a multiple of 8, I = 64), fewer slabs than CTAs per tile type,
    ragged last slab; checked against an fp64 contraction (rtol 4e-6: the chained tensor-core accumulation carries a
    flat truncation bias of about -2e-6, see lstm_tc_wgrad.cu)."""
    from hy2dl_b200 import _lib, ops
    dev = torch.device("cuda:0")
    IP = (I + 7) // 8 * 8
    rng = np.random.default_rng(B + I)
    dG = rng.normal(0, 1, (T, B, 4 * H)).astype(np.float32)       # columns g*H + u
    h = rng.normal(0, 1, (T, B, H)).astype(np.float32)
    x = np.zeros((T, B, IP), np.float32); x[:, :, :I] = rng.normal(0, 1, (T, B, I))
    cu = lambda a: torch.tensor(a, device=dev)
    dwi = torch.full((4 * H, I), np.nan, device=dev); dwh = torch.full((4 * H, H), np.nan, device=dev)
    dbb = torch.full((4 * H,), np.nan, device=dev)
    ws = ops._ws(_lib.WS_LSTM_WGRAD, B, T, I, H, dev, _lib.PREC_FP32)
    # the gate-gradient tape of this chain is blocked: [T][ceil(B/32)][4H/8][32 sequences][8]  (RB8)
    G = (B + 31) // 32
    dG_p = np.zeros((T, G * 32, 4 * H), np.float32); dG_p[:, :B] = dG
    dG_rb8 = np.ascontiguousarray(dG_p.reshape(T, G, 32, 4 * H // 8, 8).transpose(0, 1, 3, 2, 4))
    args = [cu(a) for a in (dG_rb8, h, x)]
    _lib.call("hy2dl_lstm_wgrad", *[a.data_ptr() for a in args], B, T, I, IP, H, dwi.data_ptr(), dwh.data_ptr(),
              dbb.data_ptr(), ws.data_ptr(), ws.numel(), _lib.PREC_FP32, None, _lib.stream())
    torch.cuda.synchronize()
    d = dG.astype(np.float64)
    hprev = np.concatenate([np.zeros_like(h[:1]), h[:-1]]).astype(np.float64)
    assert_close(dwi.cpu().numpy(), np.einsum("tbm,tbi->mi", d, x[:, :, :I].astype(np.float64)), rtol=4e-6, what="dW_ih")
    assert_close(dwh.cpu().numpy(), np.einsum("tbm,tbk->mk", d, hprev), rtol=4e-6, atol=1e-30, what="dW_hh")
    assert_close(dbb.cpu().numpy(), d.sum((0, 1)), rtol=4e-6, what="db")
