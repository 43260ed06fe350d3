"""Multi-frequency LSTM (BASELINE.json config 5): one LSTM cell run over a coarse (daily) stretch followed by a fine
(hourly) stretch of the same window, with one input projection per frequency and shared recurrent weights.

PARITY UNPINNED: the reference snapshot holds no multi-frequency model (SURVEY.md section 0.5 -- README.md:10 only
announces it for the successor repository), so there is no reference code or golden vector to compare with.  The
definition below follows SURVEY.md section 8(f) rank 4 ("a composition of the recurrent kernels over a concatenated
365 d + 336 h step sequence with one input projection per frequency"); the tests check it against this project's
own fp64 restatement of the same definition (test infrastructure, `mflstm_forward`).

The per-frequency projections are realised without a new kernel: the inputs of all frequencies are laid out as
disjoint column blocks of one wide input row (zeros outside a step's own block), so the single packed
`lstm.weight_ih_l0` [4H, I_total] is the concatenation [W_freq0 | W_freq1 | W_shared | frequency-flag columns] and
the recurrent kernels (csrc/lstm_tc.cu for precision tf32x3, csrc/lstm_fp32.cu otherwise) run unchanged over the
T = sum of the stretch lengths steps."""
from __future__ import annotations

from typing import Dict, List, Union

import torch
import torch.nn as nn

from .. import ops
from .lstm import LSTM


class MFLSTM(nn.Module):
    """Hyper-parameter keys: "input_size_lstm" {freq: n_inputs} and "seq_length" {freq: n_steps}, both ordered from
    the oldest / coarsest stretch to the newest / finest (e.g. {"1D": 365, "1h": 336}); "n_shared_inputs" (the last
    k inputs of every frequency are the same variables -- static attributes -- and share their weights, default 0);
    "hidden_size", "no_of_layers" (1), "predict_last_n", "drop_out_rate", optional "precision".

    forward(x = {freq: [B, n_steps, n_inputs]}) -> {"y_hat": [B, predict_last_n, 1]} from the last steps of the
    finest stretch.  State-dict keys: lstm.weight_ih_l0 [4H, I_total], lstm.weight_hh_l0, lstm.bias_ih_l0,
    lstm.bias_hh_l0, linear.weight, linear.bias; `columns` maps every frequency to its column indices."""

    def __init__(self, hyperparameters: Dict[str, Union[int, float, str, dict]]):
        super().__init__()
        sizes: Dict[str, int] = dict(hyperparameters["input_size_lstm"])
        self.seq_length: Dict[str, int] = dict(hyperparameters["seq_length"])
        if list(sizes) != list(self.seq_length):
            raise ValueError("input_size_lstm and seq_length must list the same frequencies in the same order")
        self.frequencies: List[str] = list(sizes)
        self.n_shared = int(hyperparameters.get("n_shared_inputs", 0))
        if any(n < self.n_shared for n in sizes.values()):
            raise ValueError("n_shared_inputs exceeds a frequency's input count")
        self.hidden_size = hyperparameters["hidden_size"]
        self.predict_last_n = hyperparameters["predict_last_n"]
        if not (1 <= self.predict_last_n <= self.seq_length[self.frequencies[-1]]):
            raise ValueError("predict_last_n must lie within the finest stretch")
        # column layout: own block per frequency, then the shared block, then one flag column per frequency after
        # the first (its weight column is that frequency's bias offset)
        self.columns: Dict[str, List[int]] = {}
        col = 0
        for f in self.frequencies:
            own = sizes[f] - self.n_shared
            self.columns[f] = list(range(col, col + own))
            col += own
        shared = list(range(col, col + self.n_shared))
        col += self.n_shared
        self.flag_column: Dict[str, int] = {}
        for f in self.frequencies[1:]:
            self.flag_column[f] = col
            col += 1
        for f in self.frequencies:
            self.columns[f] += shared
        self.input_size_total = col
        self.lstm = LSTM(input_size=col, hidden_size=self.hidden_size, batch_first=True,
                         num_layers=hyperparameters["no_of_layers"], precision=hyperparameters.get("precision", "auto"))
        self.dropout = nn.Dropout(hyperparameters.get("drop_out_rate", 0.0))
        self.linear = nn.Linear(in_features=self.hidden_size, out_features=1)

    def pack(self, x: Dict[str, torch.Tensor]) -> torch.Tensor:
        """{freq: [B, n_steps, n_inputs]} -> [B, T_total, I_total] block layout (plain tensor copies)."""
        first = x[self.frequencies[0]]
        B, T = first.shape[0], sum(self.seq_length.values())
        out = first.new_zeros(B, T, self.input_size_total)
        t0 = 0
        for f in self.frequencies:
            xf, n = x[f], self.seq_length[f]
            if tuple(xf.shape) != (B, n, len(self.columns[f])):
                raise ValueError(f"x['{f}'] has shape {tuple(xf.shape)}, expected {(B, n, len(self.columns[f]))}")
            out[:, t0:t0 + n, self.columns[f]] = xf
            if f in self.flag_column:
                out[:, t0:t0 + n, self.flag_column[f]] = 1.0
            t0 += n
        return out

    def forward(self, x: Dict[str, torch.Tensor]):
        xc = self.pack(x)
        B, T, _ = xc.shape
        t0 = T - self.predict_last_n
        h_seq, _ = self.lstm.run_native(ops.pack_x(xc, self.lstm.layout_for(B)), dh_t0=t0, want_seq=True, B=B)
        if ops.is_ri32(h_seq):
            h_seq = ops.from_ri32(h_seq, B)
        out = self.dropout(h_seq[t0:].permute(1, 0, 2))     # [B, n_last, H]
        return {"y_hat": self.linear(out)}
