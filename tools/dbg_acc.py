"""Accuracy of the fp32 and tf32x3 recurrent paths against the fp64 oracle (relative to max |ref|)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from test_gpu_tc import _bwd_both
from hy2dl_b200 import _lib
dev = torch.device("cuda:0")
for (B, T, I, t0, last) in [(96, 365, 25, 183, False), (256, 365, 25, 0, True), (96, 730, 25, 365, False)]:
    out, ref, x, wg = _bwd_both(dev, B, T, I, t0, last)
    line = f"B={B} T={T}: "
    for prec, name in ((_lib.PREC_FP32, "fp32"), (_lib.PREC_TF32X3, "tf32x3")):
        errs = []
        for k in ("lstm.weight_ih_l0", "lstm.weight_hh_l0", "lstm.bias_ih_l0"):
            errs.append(np.abs(wg[prec][k] - ref[k]).max() / np.abs(ref[k]).max())
        line += f"{name}: dW_ih {errs[0]:.2e} dW_hh {errs[1]:.2e} db {errs[2]:.2e}   "
    print(line)
