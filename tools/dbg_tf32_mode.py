"""Errors of the throughput mode (precision 'tf32') against the reference goldens: python tools/dbg_tf32_mode.py"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from _util import cudalstm_weights_np, golden
from hy2dl_b200.aux_functions import nse_basin_averaged, weighted_rmse
from hy2dl_b200.modelzoo import HBV, SHM, CudaLSTM, Hybrid, UH_routing
from hy2dl_b200.synthetic import lstm_weights
dev = torch.device("cuda:0")
cu = lambda a: torch.tensor(np.asarray(a), dtype=torch.float32, device=dev)
rel = lambda a, b: float(np.max(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64))) / np.max(np.abs(b)))
for prec in ("auto", "tf32"):
    for name in ("cudalstm_c1", "cudalstm_c3_full", "cudalstm_c3_h128"):
        g = golden(name); B, T, I, H, ws = [int(v) for v in g["meta"]][:5]
        m = CudaLSTM({"input_size_lstm": I, "hidden_size": H, "no_of_layers": 1, "drop_out_rate": 0.0, "precision": prec})
        m.load_state_dict({k: torch.tensor(v) for k, v in cudalstm_weights_np(g).items()}); m = m.to(dev)
        y = m(cu(g["x_lstm"]))["y_hat"]
        nse_basin_averaged(y_sim=y, y_obs=cu(g["y_obs"]), per_basin_target_std=cu(g["basin_std"])).backward()
        print(prec, name, "y %.2e" % rel(y.detach().cpu(), g["y_hat"]), {k.split(".")[-1]: "%.1e" % rel(p.grad.cpu(), g["grad." + k]) for k, p in m.named_parameters()})
    for name in ("hybrid_shm_c2", "hybrid_hbv_c4"):
        g = golden(name); B, T, n_last, I, H, mm, ws, routing = [int(v) for v in g["meta"]]
        cls = {"shm": SHM, "hbv": HBV}[str(g["model"])]
        hp = {"input_size_lstm": I, "hidden_size": H, "no_of_layers": 1, "seq_length": T, "predict_last_n": n_last,
              "n_conceptual_models": mm, "conceptual_dynamic_parameterization": [str(s) for s in g["dynamic"]], "precision": prec}
        m = Hybrid(hp, conceptual_model=cls, routing_model=UH_routing)
        m.load_state_dict({k: torch.tensor(v) for k, v in lstm_weights(I, H, m.linear.out_features, seed=ws).items()}); m = m.to(dev)
        pred = m(x_lstm=cu(g["x_lstm"]), x_conceptual=cu(g["x_conceptual"]))
        loss = (nse_basin_averaged(y_sim=pred["y_hat"], y_obs=cu(g["y_obs"]), per_basin_target_std=cu(g["basin_std"])) if str(g["loss_kind"]) == "nse"
                else weighted_rmse(y_sim=pred["y_hat"], y_obs=cu(g["y_obs"])))
        loss.backward()
        print(prec, name, "y %.2e loss %.2e" % (rel(pred["y_hat"].detach().cpu(), g["y_hat"]), rel(loss.item(), g["loss"])),
              {k.split(".", 1)[-1]: "%.1e" % rel(p.grad.cpu(), g["grad." + k]) for k, p in m.named_parameters()})
