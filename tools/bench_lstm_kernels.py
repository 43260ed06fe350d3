"""Times the three recurrent-path kernels (fwd, bwd, wgrad) through the C ABI on random data.
usage: python tools/bench_lstm_kernels.py [--B 18944] [--T 365] [--prec tf32x3|fp32] [--iters 3]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hy2dl_b200 import _lib, ops

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=18944); ap.add_argument("--T", type=int, default=365)
ap.add_argument("--I", type=int, default=25); ap.add_argument("--H", type=int, default=64); ap.add_argument("--prec", default="tf32x3"); ap.add_argument("--iters", type=int, default=3)
a = ap.parse_args()
dev = torch.device("cuda:0")
B, T, I, H = a.B, a.T, a.I, a.H
prec = ops.PRECISIONS[a.prec]
ri = a.prec == "tf32x3"
IP = 32 if ri else (I + 7) // 8 * 8
G = (B + 31) // 32
torch.manual_seed(0)
k = 1 / float(np.sqrt(H))
w_ih = (torch.rand(4 * H, I, device=dev) * 2 - 1) * k; w_hh = (torch.rand(4 * H, H, device=dev) * 2 - 1) * k
b_ih = (torch.rand(4 * H, device=dev) * 2 - 1) * k; b_hh = (torch.rand(4 * H, device=dev) * 2 - 1) * k
seq = (lambda F: torch.empty(T, G, F // 32, 32, 32, device=dev)) if ri else (lambda F: torch.empty(T, B, F, device=dev))
x = seq(IP); x.normal_()
tape = seq if ri else (lambda F: torch.empty(T, G * 32, F, device=dev))     # RB8 tapes hold whole 32-row groups
h, c, g = seq(H), tape(H), tape(4 * H)
dh = seq(H); dh.normal_()
hl = torch.empty(B, H, device=dev)
dwi = torch.empty(4 * H, I, device=dev); dwh = torch.empty(4 * H, H, device=dev); db = torch.empty(4 * H, device=dev)
wsf = ops._ws(_lib.WS_LSTM_FWD, B, T, I, H, dev, prec); wsb = ops._ws(_lib.WS_LSTM_BWD, B, T, I, H, dev, prec)
wsw = ops._ws(_lib.WS_LSTM_WGRAD, B, T, I, H, dev, prec)
st = _lib.stream()

def fwd():
    _lib.call("hy2dl_lstm_fwd", x.data_ptr(), B, T, I, IP, H, w_ih.data_ptr(), w_hh.data_ptr(), b_ih.data_ptr(), b_hh.data_ptr(),
              h.data_ptr(), c.data_ptr(), g.data_ptr(), hl.data_ptr(), wsf.data_ptr(), wsf.numel(), prec, None, st)
def bwd():
    _lib.call("hy2dl_lstm_bwd", B, T, H, w_hh.data_ptr(), g.data_ptr(), c.data_ptr(), dh.data_ptr(), T // 2, None, g.data_ptr(),
              wsb.data_ptr(), wsb.numel(), prec, None, st)
def wgrad():
    _lib.call("hy2dl_lstm_wgrad", g.data_ptr(), h.data_ptr(), x.data_ptr(), B, T, I, IP, H, dwi.data_ptr(), dwh.data_ptr(), db.data_ptr(),
              wsw.data_ptr(), wsw.numel(), prec, None, st)

def timeit(f):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); f(); e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e)

res = {"fwd": [], "bwd": [], "wgrad": []}
for it in range(a.iters + 1):
    tf = timeit(fwd); tb = timeit(bwd); tw = timeit(wgrad)
    if it:
        res["fwd"].append(tf); res["bwd"].append(tb); res["wgrad"].append(tw)
alg = {"fwd": 4 * T * (I + 6 * H), "bwd": 4 * T * 6 * H + 4 * (T - T // 2) * H, "wgrad": 4 * T * I}
act = {"fwd": 4 * T * (IP + 6 * H), "bwd": 4 * T * (5 * H + 4 * H + H) + 4 * (T - T // 2) * H, "wgrad": 4 * T * (4 * H + H + IP)}
for k2, v in res.items():
    ms = float(np.median(v))
    print(f"{a.prec} B={B} T={T} {k2}: {ms:.3f} ms  {B / ms / 1e3:.2f} Mseq/s  algorithmic {alg[k2] * B / ms / 1e6:.0f} GB/s  "
          f"actual-traffic {act[k2] * B / ms / 1e6:.0f} GB/s")
