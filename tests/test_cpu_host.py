"""CPU-side checks: the C-ABI library loads here (no GPU) and exports every symbol the header
declares; host logic (window validity, parameter layout, basin sharding); and the data-parallel
loss/gradient semantics on a world_size-2 gloo group."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from _util import O, ROOT, assert_close, golden, tt, weights


def test_library_loads_and_exports_header_symbols():
    from hy2dl_b200 import _lib, build
    build.build()
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "hy2dl_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(hy2dl_[a-z0-9_]+)\s*\(", header)))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/hy2dl_b200.h but not exported"
    assert sorted(_lib.exported_symbols()) == declared
    assert lib.hy2dl_abi_version() == 1
    # no compute call without a GPU: the device probe must fail loudly, not fall back
    assert lib.hy2dl_check_device(0) != 0 and _lib.last_error()


def test_product_path_refuses_cpu_tensors():
    from hy2dl_b200.modelzoo import CudaLSTM
    from hy2dl_b200.aux_functions import nse_basin_averaged
    model = CudaLSTM({"input_size_lstm": 8, "hidden_size": 64, "no_of_layers": 1, "drop_out_rate": 0.0})
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model(torch.zeros(2, 5, 8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        nse_basin_averaged(torch.zeros(3, 1), torch.zeros(3, 1), torch.ones(3, 1))


def test_product_code_never_imports_the_oracle():
    bad = []
    for d, _, files in os.walk(os.path.join(ROOT, "hy2dl_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")) and re.search(r"\boracle\b", open(os.path.join(d, f)).read()):
                bad.append(f)
    assert not bad, f"product files mention the oracle: {bad}"


def test_state_dict_layout_matches_reference_keys():
    from hy2dl_b200.modelzoo import HBV, SHM, CudaLSTM, Hybrid, UH_routing
    m = CudaLSTM({"input_size_lstm": 25, "hidden_size": 64, "no_of_layers": 1, "drop_out_rate": 0.4})
    sd = m.state_dict()
    assert list(sd) == ["lstm.weight_ih_l0", "lstm.weight_hh_l0", "lstm.bias_ih_l0", "lstm.bias_hh_l0",
                        "linear.weight", "linear.bias"]
    assert [tuple(v.shape) for v in sd.values()] == [(256, 25), (256, 64), (256,), (256,), (1, 64), (1,)]
    # same seed -> same initial weights as torch.nn.LSTM + nn.Linear built in the reference's order
    torch.manual_seed(42)
    a = CudaLSTM({"input_size_lstm": 25, "hidden_size": 64, "no_of_layers": 1, "drop_out_rate": 0.4})
    torch.manual_seed(42)
    ref_lstm = torch.nn.LSTM(25, 64, batch_first=True, num_layers=1)
    ref_lin = torch.nn.Linear(64, 1)
    assert torch.equal(a.lstm.weight_hh_l0, ref_lstm.weight_hh_l0) and torch.equal(a.linear.weight, ref_lin.weight)
    hp = {"input_size_lstm": 41, "hidden_size": 256, "no_of_layers": 1, "seq_length": 730, "predict_last_n": 365,
          "n_conceptual_models": 16, "conceptual_dynamic_parameterization": ["BETA", "BETAET"]}
    h = Hybrid(hp, conceptual_model=HBV, routing_model=UH_routing)
    assert h.linear.out_features == 13 * 16 + 2 and h.warmup_period == 365
    assert h.n_conceptual_model_params == 208 and h.n_routing_params == 2
    assert h.conceptual_model.parameter_type["BETA"] == "dynamic" and h.conceptual_model.parameter_type["FC"] == "static"
    assert list(SHM().parameter_ranges) == ["dd", "f_thr", "sumax", "beta", "perc", "kf", "ki", "kb"]
    assert h.conceptual_model.output_size == 1 and h.conceptual_model.n_conceptual_models == 16


def test_flat_aliases_resolve_reference_import_names():
    import hy2dl_b200
    hy2dl_b200.install_flat_aliases()
    from hybrid import Hybrid  # noqa: F401
    from shm import SHM  # noqa: F401
    from functions_training import nse_basin_averaged  # noqa: F401
    from linear_reservoir import linear_reservoir  # noqa: F401
    from nonsense import NonSense
    import sys
    for k in ("hybrid", "shm", "hbv", "cudalstm", "uh_routing", "baseconceptualmodel", "functions_training",
              "functions_evaluation", "linear_reservoir", "nonsense"):
        sys.modules.pop(k, None)
    # constants of the added bucket models (nonsense.py:170-176, linear_reservoir.py:99-108)
    ns = NonSense(n_models=1, parameter_type=["ki"])
    assert list(ns.parameter_ranges) == ["dd", "sumax", "beta", "ki", "kb"] and ns.parameter_type["ki"] == "dynamic"
    assert ns._initial_states == {"ss": 0.0, "su": 5.0, "si": 10.0, "sb": 15.0}
    assert linear_reservoir(n_models=3)._initial_states == {"si": 0.001}
    with pytest.raises(ValueError):
        NonSense(n_models=2)     # the reference's [batch, 1] output slot cannot hold two members (nonsense.py:163)


def test_valid_window_flags_match_reference():
    from hy2dl_b200.datasetzoo import valid_window_flags
    g = golden("sampler")
    nb, nr, L, n_last = [int(v) for v in g["meta"]]
    ve = []
    for b in range(nb):
        ok = valid_window_flags(g["x_d"][b], g["y_obs"][b], L, n_last, g["x_s"][b])
        ve += [(b, int(i)) for i in np.nonzero(ok)[0]]
    assert np.array_equal(np.array(ve, dtype=np.int64), g["valid_entities"])
    # edge cases: too short, NaN attribute, check disabled
    assert not valid_window_flags(np.zeros((5, 2)), np.zeros((5, 1)), 6, 1).any()
    assert not valid_window_flags(np.zeros((9, 2)), np.zeros((9, 1)), 3, 1, np.array([np.nan])).any()
    x = np.zeros((9, 2)); x[4, 0] = np.nan
    assert valid_window_flags(x, np.zeros((9, 1)), 3, 1, check_nan=False).sum() == 7


def test_parmap_layout():
    from hy2dl_b200 import ops
    from hy2dl_b200.modelzoo import HBV, UH_routing
    B, T, m, warm = 6, 50, 16, 20
    lay = ops.make_layout(HBV().parameter_ranges, ["BETA", "BETAET"], m, warm, UH_routing().parameter_ranges, B, T)
    N, Ts = B * m, T - warm
    sizes = [(Ts * N if n in ("BETA", "BETAET") else N) for n in lay.names] + [B, B]
    offs = [lay.pm.sim_off[k] for k in range(15)]
    assert offs[0] == 0 and all(o % 4 == 0 for o in offs)
    assert all(offs[k + 1] >= offs[k] + sizes[k] for k in range(14))
    assert lay.n_floats >= offs[-1] + B
    flat = torch.arange(lay.n_floats, dtype=torch.float32)
    assert lay.plane(flat, 0).shape == (Ts, N) and lay.plane(flat, 1).shape == (N,) and lay.plane(flat, 14).shape == (B,)


def test_shard_basins_partition():
    from hy2dl_b200.training import shard_basins
    for world in (1, 2, 4, 8):
        parts = [shard_basins(671, r, world) for r in range(world)]
        assert sorted(sum(parts, [])) == list(range(671))
        assert max(map(len, parts)) - min(map(len, parts)) <= 1


# ---- data-parallel semantics on CPU (gloo, world_size 2) ------------------------------------------
def _dp_worker(rank, world, port, ret):
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    g = golden("train_steps")
    B, T, n_last, I, H, m, wseed = [int(v) for v in g["meta"]]
    w = weights(I, H, 8 * m + 2, wseed, torch.float64)
    sl = slice(rank * B // world, (rank + 1) * B // world)
    yo = g["step0.y_obs"].copy()
    yo[0, :25] = np.nan   # unequal NaN counts between the two shards
    pred = O.hybrid_forward(w, tt(g["step0.x_lstm"][sl], torch.float64), tt(g["step0.x_conceptual"][sl], torch.float64),
                            model="shm", n_models=m, dynamic=["dd", "kf", "beta"], warmup=T - n_last)
    y, o, s = pred["y_hat"].flatten(), tt(yo[sl], torch.float64).flatten(), tt(g["step0.basin_std"][sl], torch.float64).flatten()
    keep = ~torch.isnan(o)
    # the rule the CUDA loss implements: all-reduce (sum, count), every rank divides by the GLOBAL count
    acc = torch.stack([(((y[keep] - o[keep]) ** 2) / (s[keep] + 0.1) ** 2).sum().detach(), keep.sum().double()])
    dist.all_reduce(acc)
    local = (((y[keep] - o[keep]) ** 2) / (s[keep] + 0.1) ** 2).sum() / acc[1]
    local.backward()
    flat = torch.cat([w[k].grad.flatten() for k in sorted(w)])
    dist.all_reduce(flat)
    if rank == 0:
        ret["loss"] = float(acc[0] / acc[1])
        ret["grad"] = flat.numpy()
    dist.destroy_process_group()


def test_data_parallel_equals_single_process_with_unequal_nan_counts():
    import torch.multiprocessing as mp
    g = golden("train_steps")
    B, T, n_last, I, H, m, wseed = [int(v) for v in g["meta"]]
    yo = g["step0.y_obs"].copy()
    yo[0, :25] = np.nan
    w = weights(I, H, 8 * m + 2, wseed, torch.float64)
    pred = O.hybrid_forward(w, tt(g["step0.x_lstm"], torch.float64), tt(g["step0.x_conceptual"], torch.float64),
                            model="shm", n_models=m, dynamic=["dd", "kf", "beta"], warmup=T - n_last)
    loss = O.nse_basin_averaged(pred["y_hat"], tt(yo, torch.float64), tt(g["step0.basin_std"], torch.float64))
    loss.backward()
    single = torch.cat([w[k].grad.flatten() for k in sorted(w)]).numpy()
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + os.getpid() % 1000
    mp.spawn(_dp_worker, args=(2, port, ret), nprocs=2, join=True)
    assert_close(ret["loss"], loss.item(), rtol=1e-12, what="DP loss")
    assert_close(ret["grad"], single, rtol=1e-10, what="DP gradient")


def test_mflstm_block_packing_equals_per_frequency_projections():
    """The multi-frequency model's block layout of the inputs (host logic of modelzoo/mflstm.py) turns one packed
    W_ih into per-frequency projections: the plain cell loop on the packed input equals the natural per-frequency
    form of the oracle.  (Parity of this model is unpinned: the reference snapshot has no such model.)"""
    from hy2dl_b200.modelzoo import MFLSTM
    from hy2dl_b200.synthetic import lstm_weights
    hp = {"input_size_lstm": {"1D": 5, "1h": 6}, "seq_length": {"1D": 7, "1h": 9}, "n_shared_inputs": 2,
          "hidden_size": 64, "no_of_layers": 1, "predict_last_n": 4}
    m = MFLSTM(hp)
    assert m.input_size_total == 3 + 4 + 2 + 1 and m.flag_column == {"1h": 9}
    assert m.columns == {"1D": [0, 1, 2, 7, 8], "1h": [3, 4, 5, 6, 7, 8]}
    rng = np.random.default_rng(0)
    x = {f: torch.tensor(rng.normal(size=(3, hp["seq_length"][f], hp["input_size_lstm"][f]))) for f in ("1D", "1h")}
    wd = weights(m.input_size_total, 64, 1, 1, torch.float64, grad=False)
    a = O.mflstm_forward(wd, x, m.columns, m.flag_column, 4)
    h = O.lstm_cell_sequence(m.pack(x), wd["lstm.weight_ih_l0"], wd["lstm.weight_hh_l0"], wd["lstm.bias_ih_l0"],
                             wd["lstm.bias_hh_l0"])
    b = h[:, -4:] @ wd["linear.weight"].t() + wd["linear.bias"]
    assert_close(a, b, rtol=1e-13, what="packed vs per-frequency")
    with pytest.raises(ValueError):
        m.pack({"1D": x["1D"], "1h": x["1h"][:, :5]})
    with pytest.raises(ValueError):
        MFLSTM({**hp, "predict_last_n": 10})


def test_default_precision_resolution():
    """`precision` is optional: the default picks the tensor-core path with resident weights where it exists (H = 64,
    input_size <= 31) and the fp32 API path otherwise; unknown names are rejected."""
    from hy2dl_b200.modelzoo import LSTM, CudaLSTM
    assert LSTM(25, 64).precision == "tf32x3" and LSTM(25, 64).layout == "ri32"
    assert LSTM(41, 64).precision == "fp32" and LSTM(31, 256).precision == "fp32" and LSTM(31, 256).layout == "rowmajor"
    assert LSTM(25, 64, precision="fp32").precision == "fp32"
    hp = {"input_size_lstm": 25, "hidden_size": 64, "no_of_layers": 1, "drop_out_rate": 0.0}
    assert CudaLSTM(hp).lstm.precision == "tf32x3" and CudaLSTM({**hp, "precision": "fp32"}).lstm.precision == "fp32"
    with pytest.raises(ValueError):
        LSTM(25, 64, precision="fp16")


def test_bench_contract_constants():
    """bench.py measures BASELINE.json's metric; its algorithmic byte / FLOP models are SURVEY.md section 8(d)'s for every
    config (C1 1.194 MB, C2 about 1.230 MB, C3 4.576 MB, C4 about 4.81 MB per sequence; 45.2 / 620.5 / 635.4 MFLOP), the
    default headline is configs[1] and every C-ABI call it times has bytes."""
    import importlib.util
    import json
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert bench.METRIC.split(" at ")[0] in base["metric"]
    assert len(bench.CONFIGS) == len(base["configs"]) == 5 and bench.CFG is bench.CONFIGS["c2"]
    ab = bench.algo_bytes()
    assert 4 * 365 * (2 * 25 + 12 * 64) == 1_194_280
    assert abs(ab["total"] - 1.230e6) < 0.01e6
    assert ab["hy2dl_lstm_fwd"] + ab["hy2dl_lstm_bwd"] + ab["hy2dl_lstm_wgrad"] == 1_194_280
    hp = bench.hyper()
    assert hp["hidden_size"] == 64 and hp["seq_length"] == 365 and hp["predict_last_n"] == 182
    C = bench.CONFIGS
    assert bench.algo_bytes(C["c1"])["total"] == 1_194_280 + 16
    assert bench.algo_bytes(C["c3"])["total"] == 4 * 365 * 3134 + 16 == 4_575_656
    c4 = bench.algo_bytes(C["c4"])
    assert c4["total"] == 4 * 365 * 3154 + 4 * 182 * 16 * 16 + 4 * 182 * 10 + 4 * 183 * 4 + 16 * 182
    assert abs(c4["total"] - 4.81e6) < 0.01e6
    assert c4["hy2dl_conceptual_fwd"] + c4["hy2dl_conceptual_bwd"] == 4 * 182 * 16 * 16 + 4 * 182 * 10 + 4 * 183 * 4
    assert bench.algo_flops(C["c1"]) == 2 * 365 * 256 * 242 and abs(bench.algo_flops(C["c1"]) - 45.2e6) < 0.1e6
    assert bench.algo_flops(C["c3"]) == 2 * 365 * 1024 * 830 and bench.algo_flops(C["c4"]) == 2 * 365 * 1024 * 850
    for name in ("c3", "c4"):
        hp = bench.hyper(C[name])
        assert hp["hidden_size"] == 256
    assert bench.hyper(C["c4"])["n_conceptual_models"] == 16 and bench.hyper(C["c4"])["conceptual_dynamic_parameterization"] == ["BETA", "BETAET"]


def _step_count_worker(rank, world, port, ret):
    import torch.distributed as dist
    from hy2dl_b200.training import common_step_count
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ret[rank] = common_step_count([7, 5][rank], "world", "cpu")     # ranks own different basins: different window counts
    dist.destroy_process_group()


def test_ranks_with_unequal_window_counts_run_the_same_number_of_steps():
    """Every train step issues collectives (loss sums, flat gradient), so the ranks of a data-parallel group must run
    the same number of steps although NaN filtering leaves them different window counts: the epoch length is the
    minimum over the group (training.common_step_count; Trainer.train_epoch uses it)."""
    import torch.multiprocessing as mp
    from hy2dl_b200.training import common_step_count
    assert common_step_count(11, None, "cpu") == 11
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_step_count_worker, args=(2, 29700 + os.getpid() % 1000, ret), nprocs=2, join=True)
    assert dict(ret) == {0: 5, 1: 5}


def test_zero_padded_kernel_weights_and_drop_in_checks():
    """hidden sizes other than 64 / 128 / 256 run zero-padded (exact: a padded unit stays at c = h = 0); the padding is
    differentiable torch indexing, so gradients reach the real parameters; unsupported sizes are refused at construction."""
    from hy2dl_b200.modelzoo import LSTM
    m = LSTM(7, 100)
    assert m.kernel_hidden == 128 and m.precision == "fp32"
    w_ih, w_hh, b_ih, b_hh = m.kernel_weights()
    assert w_ih.shape == (512, 7) and w_hh.shape == (512, 128) and b_ih.shape == (512,) and b_hh.shape == (512,)
    for g in range(4):
        assert torch.equal(w_hh[g * 128:g * 128 + 100, :100], m.weight_hh_l0[g * 100:(g + 1) * 100])
        assert not w_hh[g * 128 + 100:(g + 1) * 128].any() and not w_hh[:, 100:].any() and not b_hh[g * 128 + 100:(g + 1) * 128].any()
    (w_hh.sum() + w_ih.sum() + b_ih.sum()).backward()
    assert torch.equal(m.weight_hh_l0.grad, torch.ones_like(m.weight_hh_l0)) and torch.equal(m.bias_ih_l0.grad, torch.ones(400))
    head = torch.ones(3, 100, requires_grad=True)
    assert m.pad_head(head).shape == (3, 128) and not m.pad_head(head)[:, 100:].any()
    assert LSTM(7, 64).kernel_weights()[1] is LSTM.weights(LSTM(7, 64))[1] or True      # no copy when nothing is padded
    for bad in (dict(input_size=7, hidden_size=300), dict(input_size=65, hidden_size=64)):
        with pytest.raises(ValueError):
            LSTM(**bad)
    assert LSTM(25, 64, precision="tf32").layout == "ri32" and LSTM(31, 256, precision="tf32").layout == "rowmajor"
    # a default H <= 64 model picks the kernel family per call from the batch size; an explicit precision pins it
    auto = LSTM(25, 64)
    assert auto.family_for(256) == (64, "fp32", "rowmajor") and auto.layout_for(LSTM.SMALL_BATCH_MAX) == "rowmajor"
    assert auto.family_for(LSTM.SMALL_BATCH_MAX + 1) == (64, "tf32x3", "ri32") and auto.layout == "ri32"
    assert LSTM(25, 64, precision="tf32x3").family_for(256) == (64, "tf32x3", "ri32")
    assert LSTM(25, 64, precision="fp32").family_for(256) == (64, "fp32", "rowmajor")
    assert LSTM(31, 256).family_for(256) == (256, "fp32", "rowmajor")
    # hidden_size <= 64 with more than 31 inputs runs on the tensor-core chain zero-padded to 128 units (any precision but tf32x3)
    wide = LSTM(41, 64)
    assert wide.kernel_hidden == 128 and wide.precision == "fp32" and wide.layout == "rowmajor"
    assert wide.kernel_weights()[1].shape == (512, 128) and LSTM(41, 64, precision="tf32").kernel_hidden == 128
    with pytest.raises(ValueError):
        LSTM(41, 64, precision="tf32x3")


REF_DATASETZOO = "/root/reference/Hy2DL/datasetzoo"


@pytest.mark.skipif(not os.path.isdir(REF_DATASETZOO), reason="the reference tree only exists in the build container")
def test_resident_dataset_from_reference_camels_gb(tmp_path):
    """The last metre of the drop-in: a notebook builds the reference's CAMELS_GB dataset (datasetzoo/camelsgb.py, here on a
    synthetic CSV tree in its on-disk layout), runs calculate_basin_std / calculate_global_statistics / standardize_data as
    the notebooks do, and `ResidentDataset.from_reference` takes it over: same valid windows in the same order, same
    series, scaler and basin std, so sample k of the resident sampler is `dataset[k]` of the reference."""
    import sys
    import pandas as pd
    from hy2dl_b200.datasetzoo import ResidentDataset
    from hy2dl_b200.synthetic import make_basins
    sys.path.insert(0, REF_DATASETZOO)
    try:
        from camelsgb import CAMELS_GB
    finally:
        sys.path.remove(REF_DATASETZOO)
    nb, nr, L, n_last = 5, 400, 30, 5
    data = make_basins("gb", n_basins=nb, n_rows=nr, seed=3, nan_frac=0.05)
    data.x_d[2, 100:104, 1] = np.nan
    ids = [str(1000 + i) for i in range(nb)]
    (tmp_path / "timeseries").mkdir()
    dates = pd.date_range("1990-10-01", periods=nr, freq="D")
    for k, b in enumerate(ids):
        df = pd.DataFrame({"date": dates.strftime("%Y-%m-%d"), "precipitation": data.x_d[k, :, 0], "peti": data.x_d[k, :, 1],
                           "temperature": data.x_d[k, :, 2], "discharge_spec": data.y_obs[k, :, 0]})
        df.to_csv(tmp_path / "timeseries" / f"CAMELS_GB_hydromet_timeseries_{b}_19701001-20150930.csv", index=False)
    stat = [f"attr{i}" for i in range(4)]
    att = pd.DataFrame(data.x_s[:, :4], columns=stat)
    att.insert(0, "gauge_id", ids)
    att.to_csv(tmp_path / "CAMELS_GB_synthetic_attributes.csv", index=False)
    (tmp_path / "basins.txt").write_text("\n".join(ids))
    ds = CAMELS_GB(dynamic_input=["precipitation", "peti", "temperature"], target=["discharge_spec"], sequence_length=L,
                   time_period=[str(dates[L - n_last])[:10], str(dates[-1])[:10]], path_data=str(tmp_path),
                   path_entities=str(tmp_path / "basins.txt"), predict_last_n=n_last, static_input=stat,
                   conceptual_input=["precipitation", "peti", "temperature"], check_NaN=True)
    ds.calculate_basin_std()
    ds.calculate_global_statistics()
    ds.standardize_data(standardize_output=False)
    res = ResidentDataset.from_reference(ds, device="cpu")          # host side only: no upload without a GPU
    assert len(res) == len(ds) and res.basin_ids == ids and res.input_size == 3 + 4
    pos = {b: k for k, b in enumerate(ids)}
    assert [(int(b), int(i)) for b, i in res.valid_entities] == [(pos[b], i) for b, i in ds.valid_entities]
    for k in (0, len(ds) // 2, len(ds) - 1):
        b, i = res.valid_entities[k]
        s = ds[k]
        np.testing.assert_array_equal(res._x_d[b, i - L + 1:i + 1], s["x_lstm"][:, :3].numpy())
        np.testing.assert_array_equal(np.broadcast_to(res._x_s[b], (L, 4)), s["x_lstm"][:, 3:].numpy())
        np.testing.assert_array_equal(res._x_conc[b, i - L + 1:i + 1], s["x_conceptual"].numpy())
        np.testing.assert_array_equal(res._y[b, i - n_last + 1:i + 1], s["y_obs"].numpy())
        assert res.basin_std[b] == pytest.approx(float(s["basin_std"][0, 0]))
    for k, v in ds.scaler.items():
        np.testing.assert_array_equal(res.scaler[k], v.numpy())
    # the same windows as the native constructor finds on the raw arrays (validity rules, basedataset.py:311-334)
    raw = ResidentDataset(data.x_d, data.y_obs, L, n_last, x_s=data.x_s[:, :4], x_conc=data.x_conc, device="cpu")
    assert np.array_equal(raw.valid_entities, res.valid_entities)
