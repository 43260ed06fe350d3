import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
from _util import O, tt
from hy2dl_b200.modelzoo import LSTM
from hy2dl_b200.synthetic import lstm_weights
dev = torch.device("cuda:0")
B, T, I, H = 16, int(sys.argv[2]), 31, int(sys.argv[1])
rng = np.random.default_rng(3)
x = rng.normal(0, 1, (B, T, I)).astype(np.float32)
gh = rng.normal(0, 1, (B, T, H)).astype(np.float32)
w = lstm_weights(I, H, 1, seed=9)
lstm = LSTM(I, H, precision="fp32").to(dev)
lstm.load_state_dict({k[5:]: torch.tensor(v) for k, v in w.items() if k.startswith("lstm.")})
out, _ = lstm(torch.tensor(x, device=dev))
(out * torch.tensor(gh, device=dev)).sum().backward()
wd = {k: tt(v, torch.float64, True) for k, v in w.items() if k.startswith("lstm.")}
ref = O.lstm_cell_sequence(tt(x, torch.float64), wd["lstm.weight_ih_l0"], wd["lstm.weight_hh_l0"], wd["lstm.bias_ih_l0"], wd["lstm.bias_hh_l0"])
(ref * tt(gh, torch.float64)).sum().backward()
rel = lambda a, b: float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
print(f"H={H} T={T} fcomp={os.environ.get('HY2DL_RM_COMP','default')} bcomp={os.environ.get('HY2DL_RM_BCOMP','default')} tc={os.environ.get('HY2DL_FP32_TC','1')}",
      "h %.2e" % rel(out.detach().cpu().numpy(), ref.detach().numpy()),
      {k: "%.2e" % rel(p.grad.cpu().numpy(), wd["lstm." + k].grad.numpy()) for k, p in lstm.named_parameters()})
