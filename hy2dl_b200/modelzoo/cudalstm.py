"""CudaLSTM: sequence-to-one LSTM regressor (drop-in for Hy2DL/modelzoo/cudalstm.py:16-55)."""
from __future__ import annotations

from typing import Dict, Union

import torch
import torch.nn as nn

from .. import ops
from .lstm import LSTM


class CudaLSTM(nn.Module):
    """LSTM -> last hidden state -> Dropout -> Linear(H, 1); returns {"y_hat": [B, 1]}.

    Hyper-parameter keys read: input_size_lstm, hidden_size, no_of_layers, drop_out_rate
    (cudalstm.py:18-26); optional extra key "precision" ("auto" default | "fp32" | "tf32x3", see LSTM).
    State-dict keys: lstm.weight_ih_l0, lstm.weight_hh_l0, lstm.bias_ih_l0, lstm.bias_hh_l0,
    linear.weight, linear.bias.
    """

    def __init__(self, hyperparameters: Dict[str, Union[int, float, str, dict]]):
        super().__init__()
        self.input_size_lstm = hyperparameters["input_size_lstm"]
        self.hidden_size = hyperparameters["hidden_size"]
        self.num_layers = hyperparameters["no_of_layers"]
        self.lstm = LSTM(input_size=self.input_size_lstm, hidden_size=self.hidden_size, batch_first=True,
                         num_layers=self.num_layers, precision=hyperparameters.get("precision", "auto"))
        self.dropout = torch.nn.Dropout(hyperparameters["drop_out_rate"])
        self.linear = nn.Linear(in_features=self.hidden_size, out_features=1)

    def forward(self, x: torch.Tensor):
        """x: [batch_size, time_steps, input_size_lstm] on the GPU."""
        return self.forward_native(ops.pack_x(x, self.lstm.layout_for(x.shape[0])), B=x.shape[0])

    def forward_native(self, xT: torch.Tensor, B: int = None):
        """xT: [T, B, IP] (or RI32 with B given) as written by the GPU-resident sampler."""
        _, h_last = self.lstm.run_native(xT, want_seq=False, B=B)
        out = self.dropout(h_last[:, :self.hidden_size])  # PyTorch's op on [B, H]: RNG semantics stay PyTorch's
        return {"y_hat": self.linear(out)}
