"""Gradient / output error of the H = 128 / 256 chain on a golden CudaLSTM case, vs the reference golden (fp32) and vs
the fp64 oracle.  Environment switches (HY2DL_FP32_TC, HY2DL_RM_COMP, HY2DL_RM_BCOMP, ...) are read once per process:
run one process per variant.   python tools/dbg_c3_parity.py [case] [precision]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _util import O, cudalstm_weights_np, golden, tt  # noqa: E402

from hy2dl_b200.aux_functions import nse_basin_averaged  # noqa: E402
from hy2dl_b200.modelzoo import CudaLSTM  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cudalstm_c3_full"
precision = sys.argv[2] if len(sys.argv) > 2 else "auto"
g = golden(name)
B, T, I, H, wseed = [int(v) for v in g["meta"]][:5]
dev = torch.device("cuda:0")
w = cudalstm_weights_np(g)
model = CudaLSTM({"input_size_lstm": I, "hidden_size": H, "no_of_layers": 1, "drop_out_rate": 0.0, "precision": precision})
model.load_state_dict({k: torch.tensor(v) for k, v in w.items()})
model = model.to(dev)
cu = lambda a: torch.tensor(np.asarray(a), dtype=torch.float32, device=dev)
y = model(cu(g["x_lstm"]))["y_hat"]
loss = nse_basin_averaged(y_sim=y, y_obs=cu(g["y_obs"]), per_basin_target_std=cu(g["basin_std"]))
loss.backward()
wd = {k: tt(v, torch.float64, True) for k, v in w.items()}
yr = O.cudalstm_forward(wd, tt(g["x_lstm"], torch.float64))
O.nse_basin_averaged(yr, tt(g["y_obs"], torch.float64), tt(g["basin_std"], torch.float64)).backward()
rel = lambda a, b: float(np.max(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64))) / np.max(np.abs(b)))
env = {k: v for k, v in os.environ.items() if k.startswith("HY2DL_")}
out = {"case": name, "env": env, "y_vs_gold": rel(y.detach().cpu().numpy(), g["y_hat"]), "y_vs_f64": rel(y.detach().cpu().numpy(), yr.detach().numpy())}
for k, p in model.named_parameters():
    got = p.grad.detach().cpu().numpy()
    out[k] = ("%.2e" % rel(got, g["grad." + k]), "%.2e" % rel(got, wd[k].grad.numpy()), "gold_vs_f64 %.2e" % rel(g["grad." + k], wd[k].grad.numpy()))
print(out)
