"""Train-step throughput of the other BASELINE.json configs (C1, C3, C4 of SURVEY.md section 8) on one GPU, next to
C2 which bench.py reports.  Prints one JSON line per (config, batch) with the time of every C-ABI call inside a step.

    python tools/bench_configs.py [--configs c1,c3,c4] [--steps 5]

Not part of the bench contract; a tool for DESIGN.md section 6 / profiles."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CONFIGS = {
    # name: (synthetic shape, model kind, H, members, dynamic parameters, loss, n_last, precision, batches)
    "c1": ("gb", "cudalstm", 64, 0, [], "nse", 1, "tf32x3", (256, 18944)),
    "c1_fp32": ("gb", "cudalstm", 64, 0, [], "nse", 1, "fp32", (256, 9472)),
    "c2": ("gb", "shm", 64, 1, ["dd", "f_thr", "sumax", "beta", "perc", "kf", "ki", "kb"], "nse", 182, "tf32x3",
           (256, 18944)),
    "c3": ("us_lstm", "cudalstm", 256, 0, [], "nse", 1, "fp32", (256, 9472, 18944)),
    "c3_h128": ("us_lstm", "cudalstm", 128, 0, [], "nse", 1, "fp32", (256, 18944)),   # the notebook's own hidden size
    "c4": ("us", "hbv", 256, 16, ["BETA", "BETAET"], "wrmse", 182, "fp32", (256, 9472, 18944)),
    "nonsense": ("gb", "nonsense", 64, 1, ["dd", "sumax", "beta", "ki", "kb"], "nse", 182, "tf32x3", (256, 18944)),
}


def run_c5(steps, warmup, batches=(256, 9472)):
    """Multi-frequency LSTM, 365 daily + 336 hourly steps (parity unpinned, see modelzoo/mflstm.py): fwd + NSE loss on
    the last 24 hours + bwd + fused clip/Adam through the module API on device-resident random windows (the resident
    sampler has no multi-frequency gather yet, so the input block packing is inside the timed step)."""
    from hy2dl_b200.aux_functions import nse_basin_averaged
    from hy2dl_b200.modelzoo import MFLSTM
    from hy2dl_b200.training import FusedClipAdam
    dev = torch.device("cuda:0")
    hp = {"input_size_lstm": {"1D": 25, "1h": 25}, "seq_length": {"1D": 365, "1h": 336}, "n_shared_inputs": 22,
          "hidden_size": 64, "no_of_layers": 1, "predict_last_n": 24, "drop_out_rate": 0.0, "precision": "tf32x3"}
    model = MFLSTM(hp).to(dev)
    opt = FusedClipAdam(model.parameters(), lr=1e-3, max_norm=1.0)
    for B in batches:
        x = {f: torch.randn(B, hp["seq_length"][f], 25, device=dev) for f in ("1D", "1h")}
        yo, sd = torch.rand(B, 24, 1, device=dev), torch.ones(B, 24, 1, device=dev)

        def step():
            opt.zero_grad()
            loss = nse_basin_averaged(y_sim=model(x)["y_hat"], y_obs=yo, per_basin_target_std=sd)
            loss.backward()
            opt.step()
        for _ in range(warmup):
            step()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); t0.record()
        for _ in range(steps):
            step()
        t1.record(); torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / steps
        print(json.dumps({"config": "c5", "I": model.input_size_total, "H": 64, "T": 701, "precision": "tf32x3", "batch": B,
                          "ms_per_step": round(ms, 3), "seq_per_s": round(B / ms * 1e3, 1),
                          "note": "parity unpinned (no reference model); API path incl. input block packing"}), flush=True)


def run(name, steps, warmup, n_basins):
    from hy2dl_b200 import _lib
    from hy2dl_b200.datasetzoo import ResidentDataset
    from hy2dl_b200.modelzoo import HBV, SHM, CudaLSTM, Hybrid, NonSense, UH_routing
    from hy2dl_b200.synthetic import make_basins
    from hy2dl_b200.training import Trainer
    import hy2dl_b200.datasetzoo.resident as res_mod
    import hy2dl_b200.ops as ops_mod
    import hy2dl_b200.training as tr_mod

    shape, kind, H, m, dyn, loss, n_last, precision, batches = CONFIGS[name]
    dev = torch.device("cuda:0")
    T = 365
    data = make_basins(shape, n_basins=n_basins, seed=42)
    ds = ResidentDataset(data.x_d, data.y_obs, T, n_last, x_s=data.x_s, x_conc=data.x_conc, device=dev)
    ds.calculate_basin_std(); ds.calculate_global_statistics(); ds.standardize_data(standardize_output=False)
    I = ds.input_size
    if kind == "cudalstm":
        model = CudaLSTM({"input_size_lstm": I, "hidden_size": H, "no_of_layers": 1, "drop_out_rate": 0.0,
                          "precision": precision})
    else:
        hp = {"input_size_lstm": I, "hidden_size": H, "no_of_layers": 1, "seq_length": T, "predict_last_n": n_last,
              "n_conceptual_models": m, "conceptual_dynamic_parameterization": dyn, "precision": precision}
        model = Hybrid(hp, conceptual_model={"shm": SHM, "hbv": HBV, "nonsense": NonSense}[kind], routing_model=UH_routing)
    model = model.to(dev)
    trainer = Trainer(model, ds, loss=loss, lr=1e-3, max_norm=1.0)
    gen = torch.Generator(device=dev); gen.manual_seed(1)

    orig, ev, on = _lib.call, {}, {"on": False}

    def timed(fn, *a):
        if not on["on"]:
            return orig(fn, *a)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); orig(fn, *a); e.record()
        ev.setdefault(fn, []).append((s, e))

    for mod in (ops_mod, tr_mod, res_mod, _lib):
        mod.call = timed
    try:
        for B in batches:
            ev.clear()
            draw = lambda: torch.randint(0, len(ds), (B,), device=dev, generator=gen)
            on["on"] = False
            for _ in range(warmup):
                trainer.step(draw())
            torch.cuda.synchronize()
            on["on"] = True
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            for _ in range(steps):
                trainer.step(draw())
            t1.record(); torch.cuda.synchronize()
            ms = t0.elapsed_time(t1) / steps
            calls = {k: round(sum(s.elapsed_time(e) for s, e in v) / steps, 3) for k, v in ev.items()}
            print(json.dumps({"config": name, "I": I, "H": H, "precision": precision, "batch": B, "ms_per_step": round(ms, 3),
                              "seq_per_s": round(B / ms * 1e3, 1),
                              "calls_ms": dict(sorted(calls.items(), key=lambda kv: -kv[1]))}), flush=True)
    finally:
        for mod in (ops_mod, tr_mod, res_mod, _lib):
            mod.call = orig


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="c1,c3,c4")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--basins", type=int, default=64, help="synthetic basins resident on the GPU (throughput does not depend on it)")
    a = ap.parse_args()
    for c in a.configs.split(","):
        if c == "c5":
            run_c5(a.steps, a.warmup)
        else:
            run(c, a.steps, a.warmup, a.basins)
