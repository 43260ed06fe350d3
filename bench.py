#!/usr/bin/env python
"""Benchmark of the hot path: train sequences/s of the hybrid LSTM(64)+SHM(+UH) chain, seq 365,
on synthetic CAMELS-GB-shaped basins (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = GPU-resident window gather -> LSTM fwd -> head -> SHM warm-up + simulation -> routing
-> masked NSE loss -> adjoints -> BPTT -> weight gradients -> (all-reduce) -> clip + Adam.
Prints ONE JSON line (rank 0).  `value` is device-timed with inputs resident in HBM; `e2e` runs the
same step through the public nn.Module API from pinned HOST tensors, copies in the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "train sequences/sec (seq 365)"

# ------------------------------------------------------------------------------------------------
# BASELINE.json configs (SURVEY.md section 8 shorthand C1..C5).  `python bench.py` measures C2 (configs[1], the
# configuration the metric is quoted on for one GPU) as the headline line and adds a `configs` object with a short
# run of every other config; `--config cN` makes config N the headline of the line instead.
# ------------------------------------------------------------------------------------------------
SHM_DYN = ["dd", "f_thr", "sumax", "beta", "perc", "kf", "ki", "kb"]
CONFIGS = {
    # kind: cudalstm | hybrid | mflstm; shape: hy2dl_b200.synthetic.SHAPES key; D / S / m / nc: dynamic parameters, storages,
    # members, conceptual forcing columns (byte model of SURVEY.md 8d)
    "c1": dict(kind="cudalstm", shape="gb", I=25, H=64, T=365, n_last=1, loss="nse", seed=42,
               workload="cudaLSTM hidden=64, synthetic CAMELS-GB shape (671 basins, 3 forcings + 22 static attributes), seq 365, "
                        "seq-to-one, nse_basin_averaged, clip 1 + Adam 1e-3"),
    "c2": dict(kind="hybrid", shape="gb", I=25, H=64, T=365, n_last=182, loss="nse", seed=42, bucket="shm", m=1, dynamic=SHM_DYN,
               D=8, S=5, nc=3,
               workload="hybrid LSTM(64)+SHM(8 dynamic params, 1 member)+UH routing, synthetic CAMELS-GB shape (671 basins), "
                        "seq 365 = 183 warm-up + 182 simulated, nse_basin_averaged, clip 1 + Adam 1e-3"),
    "c3": dict(kind="cudalstm", shape="us_lstm", I=31, H=256, T=365, n_last=1, loss="nse", seed=42,
               workload="cudaLSTM hidden=256, synthetic CAMELS-US shape (531 basins, 5 forcings + 26 static attributes), seq 365, "
                        "seq-to-one, nse_basin_averaged, clip 1 + Adam 1e-3"),
    "c4": dict(kind="hybrid", shape="us", I=41, H=256, T=365, n_last=182, loss="wrmse", seed=111111, bucket="hbv", m=16,
               dynamic=["BETA", "BETAET"], D=2, S=5, nc=4,
               workload="hybrid LSTM(256)+HBV(16 members, BETA/BETAET dynamic, 11 static params)+UH routing, synthetic CAMELS-US "
                        "shape (531 basins, 6 forcings + 35 static attributes), seq 365 = 183 warm-up + 182 simulated, weighted_rmse, "
                        "clip 1 + Adam 1e-3"),
    "c5": dict(kind="mflstm", shape="gb", I=27, H=64, T=701, n_last=24, loss="nse", seed=42,
               workload="multi-frequency LSTM(64): 365 daily + 336 hourly steps (3 forcings + 22 shared static attributes per "
                        "frequency), last 24 hours predicted, nse_basin_averaged; PARITY UNPINNED (no such model in the reference "
                        "snapshot, SURVEY.md 0.5)"),
}
DEFAULT_BATCH = 18944      # saturating: one 128-row tile per SM (148 SMs)
CFG = CONFIGS["c2"]        # the headline config of this process (set by --config)
PRECISION = "auto"

# names kept for tests / tools that import them
T_SEQ, N_LAST, HID = CONFIGS["c2"]["T"], CONFIGS["c2"]["n_last"], CONFIGS["c2"]["H"]
WORKLOAD = CONFIGS["c2"]["workload"]


def hyper(cfg=None):
    cfg = cfg or CFG
    hp = {"input_size_lstm": cfg["I"], "hidden_size": cfg["H"], "no_of_layers": 1, "drop_out_rate": 0.0, "precision": PRECISION}
    if cfg["kind"] == "hybrid":
        hp.update({"seq_length": cfg["T"], "predict_last_n": cfg["n_last"], "n_conceptual_models": cfg["m"],
                   "conceptual_dynamic_parameterization": list(cfg["dynamic"])})
    return hp


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def algo_bytes(cfg=None):
    """Algorithmic bytes per sequence, fp32, store-gates convention (SURVEY.md 8d / DESIGN.md): the LSTM reads x twice
    and writes + re-reads {h, c, i, f, g, o}; the bucket model per simulated step and member reads D parameters + forcings,
    writes S storages + outflow and re-reads them (+ writes D gradients) in the adjoint; the loss moves 16 B per step."""
    cfg = cfg or CFG
    T, I, H = cfg["T"], cfg["I"], cfg["H"]
    ab = {"hy2dl_lstm_fwd": 4 * T * (I + 6 * H),        # read x, write h c i f g o
          "hy2dl_lstm_bwd": 4 * T * 6 * H,              # re-read the tape
          "hy2dl_lstm_wgrad": 4 * T * I}                # re-read x (dG / h re-reads are implementation traffic)
    total = 4 * T * (2 * I + 12 * H)
    if cfg["kind"] == "hybrid":
        Ts, Tw, D, S, m, nc = cfg["n_last"], T - cfg["n_last"], cfg["D"], cfg["S"], cfg["m"], cfg["nc"]
        ab["hy2dl_conceptual_fwd"] = 4 * Ts * m * (D + S) + 4 * Ts * (nc + 1) + 4 * Tw * nc
        ab["hy2dl_conceptual_bwd"] = 4 * Ts * m * (2 * D + S) + 4 * Ts * (nc + 1)
        total += 4 * Ts * m * (3 * D + 2 * S) + 4 * Ts * (2 * nc + 2) + 4 * Tw * nc + 16 * Ts
    else:
        total += 16 * cfg["n_last"]
    ab["total"] = total
    return ab


def algo_flops(cfg=None):
    """Algorithmic FLOPs per sequence of the LSTM train step: fwd 2T*4H(I+H) + bwd 2T*4H(I+2H)."""
    cfg = cfg or CFG
    return 2 * cfg["T"] * 4 * cfg["H"] * (2 * cfg["I"] + 3 * cfg["H"])


def resolved_precision(cfg):
    if PRECISION != "auto":
        return PRECISION
    return "tf32x3" if (cfg["H"] <= 64 and cfg["I"] <= 31) else "fp32"


PRECISION_TEXT = {"tf32x3": "fp32-class: tcgen05 kind::tf32, 3-term operand split, fp32 accumulate, weights resident on chip",
                  "fp32": "fp32-class: H >= 128 tcgen05 kind::tf32 3-term split (weights streamed from L2; unit-split kernels with resident weight "
                          "slices at batches up to ~10 k); H = 64 CUDA-core FMA",
                  "tf32": "THROUGHPUT MODE, not the parity path: single-pass tcgen05 kind::tf32 (10-bit operands, fp32 accumulate), "
                          "stated tolerance 2e-3 (discharge) / 2e-2 (gradients)"}


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe).  Primary source: NVML
    polled in-process every 20 ms (same counters nvidia-smi prints, no process start-up delay, so short timed regions
    still get samples); fallback: the recipe's `nvidia-smi --query-gpu ... -lms 50` child process."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index):
        self.rows, self.proc, self.index, self.source = [], None, index, None
        self.sm, self.mx, self.reasons, self.stop = [], [], set(), threading.Event()

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            return pynvml, pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
        except Exception:
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.index)

    def _poll(self, nv, h):
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        while not self.stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                self.mx.append(float(mx))
                mask = int(get_reasons(h))
                self.reasons |= {n for n, b in self.BITS.items() if mask & b}
            except Exception:
                pass
            self.stop.wait(0.02)

    def __enter__(self):
        try:
            nv, h = self._nvml_handle()
            self.source = "nvml"
            self.th = threading.Thread(target=self._poll, args=(nv, h), daemon=True)
            self.th.start()
            return self
        except Exception:
            self.source = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        self.stop.set()
        if self.proc:
            self.proc.terminate()
        if self.source:
            self.th.join(timeout=2)

    def summary(self):
        sm, mx, reasons = list(self.sm), list(self.mx), set(self.reasons)
        if self.source == "nvidia-smi":
            sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
            mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            reasons = {n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")}
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": self.source}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port (torch fp32 on the host, nn.LSTM via oneDNN)
# ------------------------------------------------------------------------------------------------
def _port_batch(cfg, batch, seed, device="cpu"):
    """Synthetic batch of the config's shape in the reference's [B, T, ...] layout (host tensors)."""
    from hy2dl_b200.synthetic import make_basins
    T, n_last = cfg["T"], cfg["n_last"]
    data = make_basins(cfg["shape"], n_basins=16, n_rows=T + 64, seed=seed)
    rng = np.random.default_rng(seed)
    bi = rng.integers(0, 16, batch)
    e = rng.integers(T - 1, data.x_d.shape[1], batch)
    xd = np.stack([data.x_d[b, i - T + 1:i + 1] for b, i in zip(bi, e)])
    xs = np.repeat(data.x_s[bi][:, None, :], T, axis=1)
    xl = np.concatenate([(xd - xd.mean((0, 1))) / (xd.std((0, 1)) + 1e-6), xs], axis=2).astype(np.float32)[:, :, :cfg["I"]]
    out = {"x_lstm": xl, "y_obs": np.nan_to_num(np.stack([data.y_obs[b, i - n_last + 1:i + 1] for b, i in zip(bi, e)]), nan=1.0),
           "basin_std": np.full((batch, n_last, 1), 1.0, dtype=np.float32)}
    if cfg["kind"] == "hybrid":
        out["x_conceptual"] = np.stack([data.x_conc[b, i - T + 1:i + 1] for b, i in zip(bi, e)])
    return {k: torch.tensor(np.ascontiguousarray(v, dtype=np.float32), device=device) for k, v in out.items()}


def port_train_steps(cfg, steps, warmup, batch=256, seed=42, device="cpu"):
    """Reference-style train step with the oracle port (same torch ops as the reference: _VF.lstm -- oneDNN on the host,
    cuDNN on a GPU --, per-step elementwise bucket loop, autograd, clip_grad_norm_, Adam).  device="cpu" is the reference
    arm / cpu_baseline (the reference itself is pure Python under /root/reference and cannot travel to the GPU box, hence
    kind = "port"); device="cuda:0" is the stated library baseline (`baseline_gpu_library`).  Returns seq/s and details."""
    from oracle import hy2dl_oracle as O
    from hy2dl_b200.synthetic import lstm_weights
    cores = os.cpu_count() or 1
    if device == "cpu":
        torch.set_num_threads(cores)
    n_out = 1
    if cfg["kind"] == "hybrid":
        n_out = len(O.MODEL_TABLE[cfg["bucket"]][0]) * cfg["m"] + 2
    elif cfg["kind"] == "mflstm":
        raise RuntimeError("the multi-frequency LSTM has no reference implementation (parity unpinned): no reference arm")
    w = {k: torch.tensor(v, device=device, requires_grad=True) for k, v in lstm_weights(cfg["I"], cfg["H"], n_out, seed=seed).items()}
    keys = list(w)
    opt = torch.optim.Adam([w[k] for k in keys], lr=1e-3)
    batches = [_port_batch(cfg, batch, seed + i, device) for i in range(2)]
    sync = (lambda: torch.cuda.synchronize()) if device != "cpu" else (lambda: None)
    times = []
    for s_ in range(warmup + steps):
        b = batches[s_ % 2]
        sync()
        t0 = time.perf_counter()
        opt.zero_grad()
        if cfg["kind"] == "hybrid":
            y = O.hybrid_forward(w, b["x_lstm"], b["x_conceptual"], model=cfg["bucket"], n_models=cfg["m"], dynamic=list(cfg["dynamic"]),
                                 warmup=cfg["T"] - cfg["n_last"], library=True)["y_hat"]
        else:
            y = O.cudalstm_forward(w, b["x_lstm"], library=True)
        yo = b["y_obs"] if cfg["kind"] == "hybrid" else b["y_obs"][:, 0]
        sd = b["basin_std"] if cfg["kind"] == "hybrid" else b["basin_std"][:, 0]
        loss = O.nse_basin_averaged(y, yo, sd) if cfg["loss"] == "nse" else O.weighted_rmse(y, yo)
        loss.backward()
        torch.nn.utils.clip_grad_norm_([w[k] for k in keys], 1)
        opt.step()
        loss.item()
        sync()
        if s_ >= warmup:
            times.append(time.perf_counter() - t0)
    per_step = float(np.mean(times))
    where = f"CPU fp32, {cores} threads" if device == "cpu" else "cuda:0 (cuDNN LSTM + eager ATen bucket loop)"
    return {"value": batch / per_step, "unit": "sequences/s", "cores": cores, "kind": "port",
            "sample": f"{steps} train steps of batch {batch} after {warmup} warm-up (oracle port, torch {torch.__version__} {where})",
            "ms_per_step": per_step * 1e3, "batch": batch}


def run_gpu_library(args):
    """One (config, batch) of the stated library baseline on cuda:0 (child process of the default bench line)."""
    r = port_train_steps(CFG, max(1, min(args.steps, 3)), 1, batch=args.batch, device="cuda:0")
    print(json.dumps({"value": r["value"], "unit": "sequences/s", "ms_per_step": r["ms_per_step"], "sample": r["sample"]}))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if CFG["kind"] == "mflstm":
        print(json.dumps({"impl": "reference", "unavailable": "the multi-frequency LSTM is not in the reference snapshot (parity unpinned)"}))
        return
    heavy = CFG["H"] >= 128
    steps = max(1, min(args.steps, 2 if heavy else 5))
    r = port_train_steps(CFG, steps, 1)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "sequences/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": 1, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": CFG["workload"], "name": args.config, "batch_per_step": 256,
                       "note": "oracle port of the reference's CPU path at the reference's own batch of 256 (our arm runs a "
                               "saturating batch per GPU); the unmodified reference measured 2.8x slower than this port in the "
                               "build container (VERDICT r1)"},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "sequences/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
class CallTimer:
    """CUDA-event timing of every C-ABI call (events on the launching stream = torch's current stream)."""
    # kernel launches behind each C-ABI call (weight packing + main kernel, finish kernels, memsets excluded)
    LAUNCHES = {"hy2dl_lstm_fwd": 2, "hy2dl_lstm_bwd": 2, "hy2dl_lstm_wgrad": 2, "hy2dl_uh_fwd": 2, "hy2dl_uh_bwd": 2,
                "hy2dl_clip_adam": 2, "hy2dl_head_map_fwd": 2, "hy2dl_head_map_bwd": 2}

    def __init__(self):
        from hy2dl_b200 import _lib
        import hy2dl_b200.datasetzoo.resident as res_mod
        import hy2dl_b200.ops as ops_mod
        import hy2dl_b200.training as tr_mod
        self.mods = (ops_mod, tr_mod, res_mod, _lib)
        self.orig = _lib.call
        self.on, self.ev, self.launches = False, {}, 0
        for m in self.mods:
            m.call = self._call

    def _call(self, name, *a):
        if not self.on:
            return self.orig(name, *a)
        self.launches += self.LAUNCHES.get(name, 1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        self.orig(name, *a)
        e.record()
        self.ev.setdefault(name, []).append((s, e))

    def reset(self):
        self.ev, self.launches = {}, 0

    def per_call_ms(self):
        return {k: float(np.mean([s.elapsed_time(e) for s, e in v])) for k, v in self.ev.items()}

    def per_step_ms(self, steps):
        return {k: float(sum(s.elapsed_time(e) for s, e in v)) / steps for k, v in self.ev.items()}

    def restore(self):
        for m in self.mods:
            m.call = self.orig


def bind_to_gpu_numa_node(local):
    """Pin this rank's host threads to the CPUs of its GPU's NUMA node BEFORE the pinned staging buffers of the e2e leg
    are allocated (first touch places them there): at N = 8 eight ranks' 802 MB/step uploads otherwise cross the socket
    interconnect (round 1: e2e efficiency 0.31 at N = 8).  Returns the node or None."""
    try:
        pr = torch.cuda.get_device_properties(local)
        pci = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{pci}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


def build_setup(cfg, dev, rank, world, group, use_graph=False):
    """Model + resident data set + trainer of one config; the data set holds this rank's basins, the scaler is global."""
    from hy2dl_b200.datasetzoo import ResidentDataset
    from hy2dl_b200.modelzoo import HBV, SHM, CudaLSTM, Hybrid, UH_routing
    from hy2dl_b200.synthetic import make_basins
    from hy2dl_b200.training import Trainer, shard_basins
    full = make_basins(cfg["shape"], seed=cfg["seed"])
    mine = shard_basins(full.x_d.shape[0], rank, world)
    ds = ResidentDataset(full.x_d, full.y_obs, cfg["T"], cfg["n_last"], x_s=full.x_s,
                         x_conc=full.x_conc if cfg["kind"] == "hybrid" else None, device=dev, basins=mine)
    ds.calculate_basin_std()
    ds.calculate_global_statistics()          # over ALL basins, so every rank standardises identically
    ds.standardize_data(standardize_output=cfg["kind"] != "hybrid")
    assert ds.input_size == cfg["I"], (ds.input_size, cfg["I"])
    torch.manual_seed(cfg["seed"])
    if cfg["kind"] == "hybrid":
        model = Hybrid(hyper(cfg), conceptual_model={"shm": SHM, "hbv": HBV}[cfg["bucket"]], routing_model=UH_routing)
    else:
        model = CudaLSTM(hyper(cfg))
    H = cfg["H"]
    model.lstm.bias_hh_l0.data[H:2 * H] = 3.0     # set_forget_gate (LSTM_CAMELS_GB.ipynb:278)
    model = model.to(dev)
    trainer = Trainer(model, ds, loss=cfg["loss"], lr=1e-3, max_norm=1.0, group=group, use_graph=use_graph)
    return model, ds, trainer


def timed_steps(step_fn, steps, warmup, timer, barrier, clock_index=None):
    """W untimed + K timed steps bracketed by barrier + synchronize; device time by CUDA events; optional clock sampling."""
    for _ in range(warmup):
        step_fn()
    barrier()
    if timer:
        timer.reset()
        timer.on = True
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clk = ClockSampler(clock_index) if clock_index is not None else None
    if clk:
        clk.__enter__()
    barrier()
    t0.record()
    out = None
    for _ in range(steps):
        out = step_fn()
    t1.record()
    barrier()
    if clk:
        clk.__exit__()
    if timer:
        timer.on = False
    return t0.elapsed_time(t1), out, (clk.summary() if clk else None)


def roofline_of(cfg, B, value_per_gpu, timer, steps, ms_total):
    """HBM roofline of the dominant kernel (algorithmic bytes / CUDA-event duration / measured peak) and of the step, plus
    the tensor-pipe ceiling of the 3xTF32 arithmetic for the configs it binds (H = 256)."""
    ab = algo_bytes(cfg)
    per_call = timer.per_call_ms()
    per_step = timer.per_step_ms(steps)
    cand = [k for k in per_call if k in ab]
    if not cand:
        return None
    top = max(cand, key=lambda k: per_step[k])
    pk, pk_kind = peaks()
    achieved = ab[top] * B / (per_call[top] * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and cfg is CONFIGS["c2"]:
        traffic = json.load(open(tpath)).get(top)
    # dense tf32 peak = half the measured bf16 figure; three MMA passes per product
    tf32_peak = pk.get("bf16_tflops_sustained", pk["bf16_tflops"]) / 2
    tensor_ceiling = tf32_peak * 1e12 / 3 / algo_flops(cfg)
    return {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
            "frac": achieved / pk["hbm_gbs"], "traffic": traffic, "peak_source": pk_kind,
            "ms_per_launch": per_call[top], "share_of_step": per_step[top] * steps / ms_total,
            "step_frac_of_hbm_roofline": ab["total"] * value_per_gpu / 1e9 / pk["hbm_gbs"],
            "hbm_ceiling_seq_per_s": pk["hbm_gbs"] * 1e9 / ab["total"],
            "tensor_ceiling_3xtf32_seq_per_s": tensor_ceiling,
            "frac_of_binding_ceiling": value_per_gpu / min(tensor_ceiling, pk["hbm_gbs"] * 1e9 / ab["total"]),
            "kernel_ms": {k: round(v, 4) for k, v in sorted(per_step.items(), key=lambda kv: -kv[1])}}


def run_c5(dev, batch, steps, warmup, barrier, group=None):
    """Multi-frequency LSTM (parity unpinned): forward + NSE on the last 24 hours + backward + fused clip/Adam through the
    module API on device-resident random windows (no multi-frequency sampler exists; input packing is inside the step)."""
    from hy2dl_b200.aux_functions import nse_basin_averaged
    from hy2dl_b200.modelzoo import MFLSTM
    from hy2dl_b200.training import FusedClipAdam
    hp = {"input_size_lstm": {"1D": 25, "1h": 25}, "seq_length": {"1D": 365, "1h": 336}, "n_shared_inputs": 22,
          "hidden_size": 64, "no_of_layers": 1, "predict_last_n": 24, "drop_out_rate": 0.0, "precision": "tf32x3"}
    model = MFLSTM(hp).to(dev)
    opt = FusedClipAdam(model.parameters(), lr=1e-3, max_norm=1.0, group=group)      # flat-gradient all-reduce under DP
    x = {f: torch.randn(batch, hp["seq_length"][f], 25, device=dev) for f in ("1D", "1h")}
    yo, sd = torch.rand(batch, 24, 1, device=dev), torch.ones(batch, 24, 1, device=dev)

    def step():
        opt.zero_grad()
        loss = nse_basin_averaged(y_sim=model(x)["y_hat"], y_obs=yo, per_basin_target_std=sd, group=group)
        loss.backward()
        opt.step()
        return loss
    ms, _, _ = timed_steps(step, steps, warmup, None, barrier)
    return {"value": batch * steps / (ms * 1e-3), "unit": "sequences/s", "batch_per_gpu": batch, "ms_per_step": ms / steps,
            "steps": steps, "precision": "tf32x3", "workload": CONFIGS["c5"]["workload"], "parity": "unpinned"}


def run_side_config(name, dev, rank, world, group, barrier, steps=3, warmup=3, batch=DEFAULT_BATCH, precision=None):
    """Short run of one non-headline config for the `configs` object (single GPU only).  precision: override for this
    run (the throughput-mode entries)."""
    global PRECISION
    if precision is not None:
        saved, PRECISION = PRECISION, precision
        try:
            return run_side_config(name, dev, rank, world, group, barrier, steps, warmup, batch)
        finally:
            PRECISION = saved
    cfg = CONFIGS[name]
    if cfg["kind"] == "mflstm":
        return run_c5(dev, 9472, steps, warmup, barrier)
    model, ds, trainer = build_setup(cfg, dev, rank, world, group)
    gen = torch.Generator(device=dev)
    gen.manual_seed(77 + rank)
    n_valid = len(ds)
    timer = CallTimer()
    try:
        ms, _, _ = timed_steps(lambda: trainer.step(torch.randint(0, n_valid, (batch,), device=dev, generator=gen)),
                               steps, warmup, timer, barrier)
        value = batch * steps / (ms * 1e-3)
        rl = roofline_of(cfg, batch, value, timer, steps, ms)
    finally:
        timer.restore()
    out = {"value": value, "unit": "sequences/s", "batch_per_gpu": batch, "ms_per_step": ms / steps, "steps": steps,
           "precision": resolved_precision(cfg), "workload": cfg["workload"], "roofline": rl,
           "algorithmic_bytes_per_sequence": algo_bytes(cfg)["total"], "algorithmic_flops_per_sequence": algo_flops(cfg)}
    # the reference's own batch of 256 (the unit-split kernels spread its 2 tiles over 8 - 32 SMs)
    # (CUDA-graph replayed like the headline config's leg: at 3 - 8 ms per step the eager launches are a tenth of it)
    from hy2dl_b200.training import Trainer
    tr2 = Trainer(model, ds, loss=cfg["loss"], group=group, use_graph=True)
    ms2, _, _ = timed_steps(lambda: tr2.step(torch.randint(0, n_valid, (256,), device=dev, generator=gen)), 10, 4, None, barrier)
    out["reference_batch"] = {"batch_per_gpu": 256, "value": 256 * 10 / (ms2 * 1e-3), "ms_per_step": ms2 / 10, "cuda_graph": True}
    del trainer, tr2, model, ds
    torch.cuda.empty_cache()
    return out


def run_ours(args):
    import torch.distributed as dist
    from hy2dl_b200 import _lib
    from hy2dl_b200.aux_functions import nse_basin_averaged, weighted_rmse
    from hy2dl_b200.training import Trainer

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_node = bind_to_gpu_numa_node(local)
    group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        group = "world"
    _lib.load()
    cfg = CFG

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    B = args.batch
    if cfg["kind"] == "mflstm":           # no sampler / trainer for C5: its own short path
        r = run_c5(dev, min(B, 9472), args.steps, args.warmup, barrier, group)
        t5 = torch.tensor([r["ms_per_step"]], device=dev)
        if world > 1:
            dist.all_reduce(t5, op=dist.ReduceOp.MAX)
            r["value"] = r["batch_per_gpu"] / (float(t5.item()) * 1e-3)
            r["ms_per_step"] = float(t5.item())
        if rank == 0:
            print(json.dumps({"metric": METRIC.replace("365", "365 d + 336 h"), "value": r["value"] * world, "unit": "sequences/s",
                              "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
                              "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                              "config": {"workload": cfg["workload"], "name": args.config, "batch_per_gpu": r["batch_per_gpu"],
                                         "parallelism": f"dp{world}, flat-gradient NCCL all-reduce" if world > 1 else "single GPU"}}))
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- data: synthetic CAMELS shape, basins sharded over ranks, resident in HBM
    model, ds, trainer = build_setup(cfg, dev, rank, world, group, use_graph=args.graph and world == 1)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    n_valid = len(ds)

    def draw(n=B):
        return torch.randint(0, n_valid, (n,), device=dev, generator=gen)

    # ---- value: resident inputs, device-timed, per-call CUDA events for the roofline object
    timer = CallTimer()
    ms, loss, clocks = timed_steps(lambda: trainer.step(draw()), args.steps, args.warmup, timer, barrier, clock_index=local)
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = B * world * args.steps / (ms * 1e-3)
    final_loss = float(loss.item())
    launches = timer.launches
    roofline = roofline_of(cfg, B, value / world, timer, args.steps, ms)
    timer.restore()
    if roofline is None:      # --graph: the step is one graph launch, no per-call events
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": value, "ms_per_step": ms / args.steps, "n_gpus": world, "cuda_graph": True}))
        return

    # ---- e2e: public nn.Module API, pinned host inputs, H2D copies + loss read-back inside the timed region
    opt = trainer.opt
    Be = args.e2e_batch or B
    hybrid = cfg["kind"] == "hybrid"
    host = []
    for _ in range(0 if args.quick else 2):
        ref = ds.gather(draw(Be)).as_reference()
        host.append({k: v.cpu().pin_memory() for k, v in ref.items()})
        del ref
    h2d = sum(v.numel() * 4 for v in host[0].values()) if host else 0
    copy_stream = torch.cuda.Stream(device=dev)      # two-deep pipeline: batch i+1 is copied while batch i computes
    compute = torch.cuda.current_stream()

    def upload(hb):
        with torch.cuda.stream(copy_stream):
            d = {k: v.to(dev, non_blocking=True) for k, v in hb.items()}
            ev_up = torch.cuda.Event()
            ev_up.record(copy_stream)
        return d, ev_up

    def e2e_step(staged, nxt):
        d, ev_up = staged
        compute.wait_event(ev_up)
        for v in d.values():
            v.record_stream(compute)
        nxt_staged = upload(nxt) if nxt is not None else None
        opt.zero_grad()
        if hybrid:
            y = model(x_lstm=d["x_lstm"], x_conceptual=d["x_conceptual"])["y_hat"]
            yo, sd = d["y_obs"], d["basin_std"]
        else:
            y = model(d["x_lstm"])["y_hat"]
            yo, sd = d["y_obs"][:, 0], d["basin_std"][:, 0]
        l = (nse_basin_averaged(y_sim=y, y_obs=yo, per_basin_target_std=sd, group=group) if cfg["loss"] == "nse"
             else weighted_rmse(y_sim=y, y_obs=yo, group=group))
        l.backward()
        opt.step()
        return l.item(), nxt_staged   # device -> host read of the step's result

    e2e_steps = 0 if args.quick else max(3, args.steps // 2)
    if e2e_steps:
        st = upload(host[0])
        for i in range(2):
            _, st = e2e_step(st, host[(i + 1) % 2])
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):                      # every step's upload is issued inside the timed region
            _, st = e2e_step(st, host[(i + 1) % 2])     # (the one consumed first was issued during warm-up; one extra is issued at the end)
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], device=dev)
    else:
        dt = torch.tensor([1.0], device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e = {"value": Be * world * e2e_steps / float(dt.item()) if e2e_steps else None, "unit": "sequences/s", "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": 4, "steps": e2e_steps, "batch_per_gpu": Be, "host_numa_node": numa_node}
    del host

    # ---- the same step through the training API (Trainer.step): the data set is resident in HBM, so a step's host
    # input is its list of sample ids (pinned, copied H2D every step) and the result read back is the loss
    e2e_res = None
    if e2e_steps:
        ids_host = [torch.randint(0, n_valid, (B,)).pin_memory() for _ in range(2)]
        for i in range(2):
            trainer.step(ids_host[i].to(dev, non_blocking=True)).item()
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            trainer.step(ids_host[i % 2].to(dev, non_blocking=True)).item()
        barrier()
        dtr = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.all_reduce(dtr, op=dist.ReduceOp.MAX)
        e2e_res = {"value": B * world * e2e_steps / float(dtr.item()), "unit": "sequences/s", "h2d_bytes_per_step": 8 * B,
                   "d2h_bytes_per_step": 4, "steps": e2e_steps, "api": "Trainer.step(ids) on the HBM-resident data set"}

    # ---- reference batch size (256 per GPU), CUDA-graph replayed: the latency-bound regime
    small = None
    if args.small_batch and not args.quick:
        model2, _, _ = build_setup(cfg, dev, rank, world, group)[0], None, None
        tr2 = Trainer(model2, ds, loss=cfg["loss"], group=group, use_graph=(world == 1))
        ids = [draw(args.small_batch) for _ in range(4)]
        for i in range(4):
            tr2.step(ids[i % 4])
        barrier()
        s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s_.record()
        n_small = 20
        for i in range(n_small):
            tr2.step(ids[i % 4])
        e_.record()
        barrier()
        small = {"batch_per_gpu": args.small_batch, "value": args.small_batch * world * n_small / (s_.elapsed_time(e_) * 1e-3),
                 "ms_per_step": s_.elapsed_time(e_) / n_small, "cuda_graph": world == 1}
        del tr2, model2

    # ---- every other BASELINE config, short runs (single GPU, default invocation only)
    configs, gpu_lib, cpu = None, None, None
    del trainer, model, ds
    torch.cuda.empty_cache()
    if world == 1 and not args.quick and not args.no_configs:
        configs = {}
        for name in CONFIGS:
            if CONFIGS[name] is cfg:
                continue
            try:
                configs[name] = run_side_config(name, dev, rank, world, group, barrier)
            except Exception as ex:   # a side config must never take the headline line down
                configs[name] = {"error": f"{type(ex).__name__}: {ex}"[:300]}
            torch.cuda.empty_cache()
        # the throughput mode (single-pass TF32; looser stated tolerance -- never the headline) on the tensor-bound configs
        for name in ("c2", "c3", "c4"):
            try:
                configs[name + "_tf32"] = run_side_config(name, dev, rank, world, group, barrier, precision="tf32")
            except Exception as ex:
                configs[name + "_tf32"] = {"error": f"{type(ex).__name__}: {ex}"[:300]}
            torch.cuda.empty_cache()
        # stated library baseline: the oracle port (same torch ops as the reference) on THIS GPU -- nn.LSTM -> cuDNN, the
        # bucket loop as eager ATen ops -- i.e. what the unmodified reference does with running_device = "gpu".  It runs in
        # a child process (`--impl gpu_library`), so that the process of our arm never calls into the cuDNN RNN path.
        gpu_lib = {}
        for name, batches in (("c1", (256, 16384)), ("c3", (256, 4096)), ("c2", (256,))):
            for b_ in batches:
                try:
                    out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "gpu_library", "--config", name,
                                          "--batch", str(b_), "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300)
                    js = [l for l in out.stdout.splitlines() if l.startswith("{")]
                    gpu_lib[f"{name}_b{b_}"] = json.loads(js[-1]) if js else {"error": (out.stderr or "no output")[-200:]}
                except Exception as ex:
                    gpu_lib[f"{name}_b{b_}"] = {"error": f"{type(ex).__name__}: {ex}"[:200]}
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not args.quick:
        r = port_train_steps(cfg, 2 if cfg["H"] >= 128 else 3, 1)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        prec = resolved_precision(cfg)
        line = {"metric": METRIC, "value": value, "unit": "sequences/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": cfg["workload"], "name": args.config, "batch_per_gpu": B, "global_batch": B * world,
                           "seq_len": cfg["T"], "hidden": cfg["H"], "precision": PRECISION_TEXT[prec],
                           "parallelism": f"dp{world} over basins, flat-gradient NCCL all-reduce" if world > 1 else "single GPU",
                           "l2_policy": "inputs larger than L2: every step gathers a new random batch; the per-step tape "
                                        f"({4 * cfg['T'] * 12 * cfg['H'] * B / 1e9:.1f} GB) exceeds the 126 MB L2"},
                "clocks": clocks, "e2e": e2e, "e2e_resident_sampler": e2e_res, "gpu_launches": launches, "roofline": roofline,
                "cpu_baseline": cpu, "reference_batch": small, "configs": configs, "baseline_gpu_library": gpu_lib,
                "final_loss": final_loss}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "gpu_library"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS), help="BASELINE.json config to measure as the headline (default c2 = configs[1])")
    ap.add_argument("--batch", type=int, default=DEFAULT_BATCH, help="sequences per GPU per step (saturating: 128-row tiles x 148 SMs)")
    ap.add_argument("--precision", default="auto", choices=["auto", "fp32", "tf32x3", "tf32"],
                    help="recurrent-kernel arithmetic (auto: what a user gets; tf32: the looser-tolerance throughput mode)")
    ap.add_argument("--e2e-batch", type=int, default=0)
    ap.add_argument("--small-batch", type=int, default=256, help="also report the reference batch size (0: skip)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the short runs of the other BASELINE configs")
    ap.add_argument("--graph", action="store_true", help="replay the whole train step as one CUDA graph (single GPU)")
    ap.add_argument("--quick", action="store_true", help="device-timed value only (profiling runs): no e2e / small batch / cpu / configs legs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    global PRECISION, CFG
    PRECISION = args.precision
    CFG = CONFIGS[args.config]
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "gpu_library":
        run_gpu_library(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
