"""Debug: decode what the tcgen05 wgrad kernel accumulates (structured inputs)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hy2dl_b200 import _lib, ops
dev = torch.device("cuda:0")
T, B, I, H = 2, 32, 8, 64
IP = 32
dG = np.zeros((T, B, 4 * H), np.float32)
for r in range(32):
    dG[1, r, r] = 1.0          # n = r picks row r
    dG[1, r, 128 + r] = 2.0    # second half
h = np.zeros((T, B, H), np.float32)
for r in range(32):
    for k in range(H):
        h[0, r, k] = r + 100 * k + 1
x = np.zeros((T, B, IP), np.float32)
for r in range(32):
    for k in range(IP):
        x[1, r, k] = 0.5 + r + 100 * k
cu = lambda a: torch.tensor(a, device=dev)
dwi = torch.full((4 * H, I), np.nan, device=dev); dwh = torch.full((4 * H, H), np.nan, device=dev)
dbb = torch.full((4 * H,), np.nan, device=dev)
ws = ops._ws(_lib.WS_LSTM_WGRAD, B, T, I, H, dev, _lib.PREC_TF32X3)
ws.zero_()
args = [ops.to_ri32(cu(a)) for a in (dG, h, x)]
_lib.call("hy2dl_lstm_wgrad", *[a.data_ptr() for a in args], B, T, I, IP, H, dwi.data_ptr(), dwh.data_ptr(),
          dbb.data_ptr(), ws.data_ptr(), ws.numel(), _lib.PREC_TF32X3, None, _lib.stream())
torch.cuda.synchronize()
ND = 112
part = ws.view(torch.float32)[: 2 * 256 * ND].view(2, 256, ND).cpu().numpy()
np.set_printoptions(linewidth=250, precision=1, suppress=True)
for c in range(2):
    print("CTA", c, "nonzero count", np.count_nonzero(part[c]))
    for n in (0, 1, 2, 31, 128, 129):
        print(" row", n, part[c, n, :8], "|x", part[c, n, 64:72], "|b", part[c, n, 96:100])
