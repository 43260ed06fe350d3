"""Forward kernel timing with parts of the tape switched off (what do the stores cost?)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hy2dl_b200 import _lib, ops
dev = torch.device("cuda:0")
B, T, I, H, IP = 18944, 365, 25, 64, 32
prec = _lib.PREC_TF32X3
G = (B + 31) // 32
torch.manual_seed(0)
k = 1 / 8.0
w_ih = (torch.rand(4 * H, I, device=dev) * 2 - 1) * k; w_hh = (torch.rand(4 * H, H, device=dev) * 2 - 1) * k
b_ih = (torch.rand(4 * H, device=dev) * 2 - 1) * k; b_hh = (torch.rand(4 * H, device=dev) * 2 - 1) * k
seq = lambda F: torch.empty(T, G, F // 32, 32, 32, device=dev)
x = seq(IP); x.normal_()
h, c, g = seq(H), seq(H), seq(4 * H)
hl = torch.empty(B, H, device=dev)
ws = ops._ws(_lib.WS_LSTM_FWD, B, T, I, H, dev, prec)
st = _lib.stream()
def run(hh, cc, gg):
    p = lambda t: None if t is None else t.data_ptr()
    _lib.call("hy2dl_lstm_fwd", x.data_ptr(), B, T, I, IP, H, w_ih.data_ptr(), w_hh.data_ptr(), b_ih.data_ptr(), b_hh.data_ptr(),
              p(hh), p(cc), p(gg), hl.data_ptr(), ws.data_ptr(), ws.numel(), prec, None, st)
for name, args in (("all", (h, c, g)), ("no gates", (h, c, None)), ("h only", (h, None, None)), ("h_last only", (None, None, None))):
    ts = []
    for it in range(3):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); run(*args); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    print(f"{name:12s} {min(ts[1:]):.3f} ms")
