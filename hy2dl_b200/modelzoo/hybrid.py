"""Hybrid: LSTM -> head -> conceptual model (-> unit-hydrograph routing).

Drop-in for Hy2DL/modelzoo/hybrid.py:24-117 with the whole chain on the hy2dl_b200 kernels:
recurrent LSTM (K2), fused Linear + sigmoid range mapping evaluated only on the steps/columns the
bucket model consumes (head.cu), per-series bucket time stepping with hand-derived adjoint (K3),
gamma routing (K3b).  Intermediate tensors stay in the time-major native layouts; the returned
dictionary holds [batch, time, ...] views of them, exactly the keys/shapes of the reference.
"""
from __future__ import annotations

from typing import Dict, Union

import torch
import torch.nn as nn

from .. import ops
from .baseconceptualmodel import BaseConceptualModel
from .lstm import LSTM


class Hybrid(nn.Module):
    """hyperparameters keys: input_size_lstm, hidden_size, no_of_layers, seq_length, predict_last_n,
    n_conceptual_models, conceptual_dynamic_parameterization (hybrid.py:32-46); optional "precision" ("auto" default, see LSTM).
    `conceptual_model` / `routing_model` are passed as classes, as in the reference (hybrid.py:47-53).
    """

    def __init__(self, hyperparameters: Dict[str, Union[int, float, str, dict]],
                 conceptual_model: BaseConceptualModel, routing_model: BaseConceptualModel = None):
        super().__init__()
        self.input_size_lstm = hyperparameters["input_size_lstm"]
        self.hidden_size = hyperparameters["hidden_size"]
        self.num_layers = hyperparameters["no_of_layers"]
        self.seq_length = hyperparameters["seq_length"]
        self.predict_last_n = hyperparameters["predict_last_n"]
        self.warmup_period = self.seq_length - self.predict_last_n

        self.lstm = LSTM(input_size=self.input_size_lstm, hidden_size=self.hidden_size, batch_first=True,
                         num_layers=self.num_layers, precision=hyperparameters.get("precision", "auto"))

        self.n_conceptual_models = hyperparameters["n_conceptual_models"]
        self.conceptual_dynamic_parameterization = hyperparameters["conceptual_dynamic_parameterization"]
        self.conceptual_model = conceptual_model(n_models=self.n_conceptual_models,
                                                 parameter_type=self.conceptual_dynamic_parameterization)
        if getattr(self.conceptual_model, "_model_id", None) is None:
            raise NotImplementedError(
                f"{type(self.conceptual_model).__name__} has no hy2dl_b200 CUDA implementation (SHM, HBV, linear_reservoir and NonSense do)")
        self.n_conceptual_model_params = len(self.conceptual_model.parameter_ranges) * self.n_conceptual_models
        if routing_model:
            self.routing_model = routing_model()
            self.n_routing_params = len(self.routing_model.parameter_ranges)
        else:
            self.n_routing_params = 0
        self.linear = nn.Linear(in_features=self.hidden_size,
                                out_features=self.n_conceptual_model_params + self.n_routing_params)
        self._layouts = {}

    # ----------------------------------------------------------------------------------------------
    def _layout(self, B: int, T: int) -> ops.ParLayout:
        key = (B, T)
        if key not in self._layouts:
            cm = self.conceptual_model
            dynamic = [k for k, v in cm.parameter_type.items() if v == "dynamic"]
            warmup = self.warmup_period  # fixed at construction, also for longer test sequences (hybrid.py:39,95-108)
            if not (1 <= warmup < T):
                raise ValueError(f"warmup_period={warmup} must satisfy 1 <= warmup < time_steps={T}")
            self._layouts[key] = ops.make_layout(
                cm.parameter_ranges, dynamic, self.n_conceptual_models, warmup,
                self.routing_model.parameter_ranges if self.n_routing_params else None, B, T)
        return self._layouts[key]

    def forward(self, x_lstm: torch.Tensor, x_conceptual: torch.Tensor):
        """x_lstm [B,T,input_size_lstm], x_conceptual [B,T,2|3|4] (GPU, fp32) -> prediction dict with
        y_hat [B,T_sim,1], parameters / internal_states {name: [B,T_sim,m]}, final_states {name: [B,m]}."""
        if x_lstm.shape[:2] != x_conceptual.shape[:2]:
            raise ValueError("x_lstm and x_conceptual must share batch and time dimensions")
        return self.forward_native(ops.pack_x(x_lstm, self.lstm.layout_for(x_lstm.shape[0])),
                                   ops.pack_forcing(self.conceptual_model._forcing(x_conceptual)))

    def forward_native(self, xT: torch.Tensor, forc: torch.Tensor):
        """Same as forward() for inputs already in native layout (xT [T,B,IP] or RI32, forc [T,3,B]), as
        the GPU-resident sampler writes them."""
        T, B = forc.shape[0], forc.shape[2]
        if xT.shape[0] != T:
            raise ValueError("xT and forc must share the time dimension")
        layout = self._layout(B, T)
        pm, cm = layout.pm, self.conceptual_model
        warmup, m, n_par = pm.warmup, self.n_conceptual_models, pm.n_par
        N, T_sim = B * m, T - warmup

        # LSTM over all T steps; only steps >= warmup receive gradient from the head
        if ops.is_ri32(xT) and self.linear.out_features <= ops.FUSED_HEAD_MAX_COLUMNS:
            # LSTM + head in one kernel: the hidden sequence is only written as the BPTT tape
            par_sim, par_warm = ops.lstm_head(xT, *self.lstm.kernel_weights(), self.lstm.pad_head(self.linear.weight),
                                              self.linear.bias, self.input_size_lstm, layout, self.lstm.precision)
        else:
            h_seq, _ = self.lstm.run_native(xT, dh_t0=warmup, want_seq=True, B=B)
            HK = h_seq.shape[-1] if not ops.is_ri32(h_seq) else None          # the family that ran (LSTM.layout_for)
            par_sim, par_warm = ops.head_map(h_seq, self.lstm.pad_head(self.linear.weight, HK), self.linear.bias, layout)

        # warm-up: frozen parameters, no gradient (hybrid.py:99-102)
        with torch.no_grad():
            w_offs = [k * N for k in range(n_par)]
            states_w, _ = ops.conceptual(forc, par_warm.view(-1), None, cm._model_id, B, m, 0, warmup, w_offs,
                                         [False] * n_par)
            init = states_w[warmup - 1]
        # simulation (hybrid.py:104-108)
        s_offs = [pm.sim_off[k] for k in range(n_par)]
        s_dyn = [bool(pm.dynamic[k]) for k in range(n_par)]
        if self.n_routing_params > 0:
            # the routing parameters leave par_sim through the same autograd node (see ConceptualFunction)
            r_off, beta_off = pm.sim_off[n_par], pm.sim_off[n_par + 1] - pm.sim_off[n_par]
            states, q, routing = ops.conceptual_routing(forc, par_sim, init, cm._model_id, B, m, warmup, T_sim, s_offs,
                                                        s_dyn, (r_off, beta_off + B))
        else:
            states, q = ops.conceptual(forc, par_sim, init, cm._model_id, B, m, warmup, T_sim, s_offs, s_dyn)

        parameters = {}
        for k, name in enumerate(layout.names):
            plane = layout.plane(par_sim, k)
            parameters[name] = (plane.view(T_sim, B, m).permute(1, 0, 2) if pm.dynamic[k]
                                else plane.view(B, 1, m).expand(B, T_sim, m))
        pred = cm._result(states, q, parameters, B, m)
        if self.n_routing_params > 0:
            pred["y_hat"] = ops.uh_route_packed(q, routing, beta_off).t().unsqueeze(2)
        return pred
