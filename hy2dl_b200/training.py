"""Train-step runtime: the body of the reference notebooks' training loops as one function.

Reference loop (experiments/LSTM_CAMELS_GB.ipynb:289-304, HybridModel_US.ipynb:276-294):
    zero_grad -> forward -> loss -> backward -> clip_grad_norm_(params, 1) -> Adam.step -> loss.item()
Here: GPU-resident window gather -> forward/backward on the hy2dl_b200 kernels -> one fused
global-norm-clip + Adam kernel over a flat parameter buffer.  No per-step host synchronisation
(the loss stays on the device), optional CUDA-graph replay of the whole step for the latency-bound
reference batch size, and data parallelism over basins with one NCCL all-reduce of the flat
gradient per step.
"""
from __future__ import annotations

from typing import Iterable, Optional

import torch
import torch.nn as nn

from . import _lib
from ._lib import call, ptr, stream
from .aux_functions.functions_training import nse_basin_averaged, weighted_rmse
from .datasetzoo.resident import Batch, ResidentDataset


class FusedClipAdam:
    """clip_grad_norm_(params, max_norm) + torch.optim.Adam(lr, betas, eps) in two kernel launches.

    The model's parameters (and their .grad) are re-pointed at slices of one flat fp32 buffer, so
    the optimiser, the gradient all-reduce and the norm each touch one contiguous array.  Parameter
    objects, names and state-dict keys are unchanged.
    """

    def __init__(self, params: Iterable[nn.Parameter], lr: float = 1e-3, max_norm: float = 1.0,
                 betas=(0.9, 0.999), eps: float = 1e-8, group=None):
        self.params = [p for p in params]
        if not self.params:
            raise ValueError("no parameters")
        dev = self.params[0].device
        for p in self.params:
            _lib.require_cuda(p.data, "parameter")
        sizes = [(p.numel() + 3) // 4 * 4 for p in self.params]  # keep every slice 16-byte aligned
        self.n = sum(sizes)
        self.flat = torch.zeros(self.n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(self.n, dtype=torch.float32, device=dev)
        off = 0
        self._grad_views = []
        for p, sz in zip(self.params, sizes):
            view = self.flat[off:off + p.numel()].view_as(p)
            view.copy_(p.data)
            p.data = view
            p.grad = self.grad[off:off + p.numel()].view_as(p)
            self._grad_views.append(p.grad)
            off += sz
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.step_count = torch.zeros((), dtype=torch.int64, device=dev)
        self.lr = torch.full((), lr, dtype=torch.float32, device=dev)
        # [0] squared norm, then the reduction's ticket and per-block partial sums (include/hy2dl_b200.h)
        self.norm_sq = torch.zeros(int(_lib.load().hy2dl_sqnorm_floats()), dtype=torch.float32, device=dev)
        self.max_norm, self.betas, self.eps, self.group = max_norm, betas, eps, group

    def set_lr(self, lr: float):
        """Scheduler hook (StepLR in LSTM_CAMELS_GB.ipynb:273,362): device scalar, graph-safe."""
        self.lr.fill_(lr)

    def zero_grad(self):
        self.grad.zero_()
        self._relink()

    def _relink(self):
        """`model.zero_grad()` (set_to_none) or an optimiser-agnostic training loop can detach a parameter's .grad from
        the flat buffer; pick such gradients up (host-side identity checks only) so that step() never updates from a
        stale buffer."""
        for p, view in zip(self.params, self._grad_views):
            if p.grad is not view:
                if p.grad is None:
                    view.zero_()
                else:
                    view.copy_(p.grad)
                p.grad = view

    def step(self):
        self._relink()
        if self.group is not None:
            import torch.distributed as dist
            dist.all_reduce(self.grad, op=dist.ReduceOp.SUM, group=None if self.group == "world" else self.group)
        self.norm_sq[:1].zero_()
        if self.max_norm > 0:
            call("hy2dl_grad_sqnorm", ptr(self.grad), self.n, ptr(self.norm_sq), stream())
        call("hy2dl_clip_adam", ptr(self.flat), ptr(self.grad), ptr(self.exp_avg), ptr(self.exp_avg_sq), self.n,
             ptr(self.norm_sq), 1.0, float(self.max_norm), ptr(self.lr), float(self.betas[0]), float(self.betas[1]),
             float(self.eps), self.step_count.data_ptr(), stream())

    def grad_norm(self) -> torch.Tensor:
        """Pre-clip global gradient norm of the last step (device scalar)."""
        return self.norm_sq[0].sqrt()


def forward_loss(model: nn.Module, batch: Batch, loss: str, group=None) -> torch.Tensor:
    """Forward + loss on a native batch; works for CudaLSTM and Hybrid."""
    if batch.forc is not None and hasattr(model, "conceptual_model"):
        pred = model.forward_native(batch.xT, batch.forc)
        y = pred["y_hat"][:, :, 0].t()       # the contiguous [T_sim, B] buffer underneath the API view
    else:
        pred = model.forward_native(batch.xT, B=batch.B)
        y = pred["y_hat"].t()                # [1, B]
    if loss == "nse":
        return nse_basin_averaged(y_sim=y, y_obs=batch.y_obs, per_basin_target_std=batch.basin_std, group=group)
    if loss == "wrmse":
        return weighted_rmse(y_sim=y, y_obs=batch.y_obs, group=group)
    raise ValueError(f"unknown loss '{loss}' (nse | wrmse)")


class Trainer:
    """One object = model + resident data set + fused optimiser; `step(ids)` is one update."""

    def __init__(self, model: nn.Module, dataset: ResidentDataset, loss: str = "nse", lr: float = 1e-3,
                 max_norm: float = 1.0, group=None, use_graph: bool = False):
        self.model, self.dataset, self.loss_kind, self.group = model, dataset, loss, group
        dataset.layout = model.lstm.layout   # the sampler writes the layout the recurrent kernels read
        self.opt = FusedClipAdam(model.parameters(), lr=lr, max_norm=max_norm, group=group)
        self.use_graph = use_graph
        self._graph = None
        self._static_ids = None
        self._static_loss = None

    def _step_eager(self, ids: torch.Tensor) -> torch.Tensor:
        self.dataset.layout = self.model.lstm.layout_for(int(ids.numel()))    # small batches of an H <= 64 model: other kernels, other layout
        batch = self.dataset.gather(ids)
        self.opt.zero_grad()
        loss = forward_loss(self.model, batch, self.loss_kind, self.group)
        loss.backward()
        self.opt.step()
        return loss.detach()

    def step(self, ids: torch.Tensor) -> torch.Tensor:
        """ids: int64 device tensor of sample indices. Returns the loss as a device scalar (no sync)."""
        if not self.use_graph:
            return self._step_eager(ids)
        if self._graph is None:
            self._static_ids = ids.clone()
            # warm-up outside capture (lazy allocations, kernel attributes); the updates it makes
            # are rolled back so that the captured step is the first real one
            o = self.opt
            saved = [t.clone() for t in (o.flat, o.exp_avg, o.exp_avg_sq, o.step_count)]
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    self._step_eager(self._static_ids)
            torch.cuda.current_stream().wait_stream(side)
            for t, v in zip((o.flat, o.exp_avg, o.exp_avg_sq, o.step_count), saved):
                t.copy_(v)
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):
                self._static_loss = self._step_eager(self._static_ids)
        self._static_ids.copy_(ids)
        self._graph.replay()
        return self._static_loss

    def train_epoch(self, batch_size: int, max_updates: Optional[int] = None,
                    generator: Optional[torch.Generator] = None) -> torch.Tensor:
        """One shuffled pass (drop_last); returns the mean loss as a device scalar."""
        n = len(self.dataset)
        order = torch.randperm(n, device=self.dataset.device, generator=generator)
        n_batches = n // batch_size
        if max_updates is not None:
            n_batches = min(n_batches, max_updates)
        n_batches = common_step_count(n_batches, self.group, self.dataset.device)
        total = torch.zeros((), dtype=torch.float32, device=self.dataset.device)
        self.model.train()
        for i in range(n_batches):
            total += self.step(order[i * batch_size:(i + 1) * batch_size])
        return total / max(n_batches, 1)


def common_step_count(n_local: int, group, device) -> int:
    """Every step issues collectives (loss partial sums, flat gradient), so all ranks of a data-parallel group must
    run the SAME number of steps.  Ranks own different basins and NaN filtering leaves them different window counts:
    the epoch length is the minimum over the group (drop-last on the shortest shard).  One small
    all-reduce per epoch; group=None: single process, nothing to do."""
    if group is None:
        return n_local
    import torch.distributed as dist
    t = torch.tensor([n_local], dtype=torch.int64, device=device if dist.get_backend() == "nccl" else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=None if group == "world" else group)
    return int(t.item())


def shard_basins(n_basins: int, rank: int, world: int):
    """Basin partition for data parallelism: rank r owns basins r, r+world, ...."""
    return list(range(rank, n_basins, world))
