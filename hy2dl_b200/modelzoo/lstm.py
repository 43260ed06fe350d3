"""Single-layer LSTM module with torch.nn.LSTM's parameter names, shapes and initialisation,
executed by the hy2dl_b200 recurrent kernels instead of cuDNN / oneDNN."""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
import torch.nn as nn

from .. import ops


class LSTM(nn.Module):
    """Replaces `nn.LSTM(input_size, hidden_size, batch_first=True, num_layers=1)` at
    modelzoo/cudalstm.py:22-24 and modelzoo/hybrid.py:41-43.

    Parameters are real `nn.Parameter`s called weight_ih_l0 [4H,I], weight_hh_l0 [4H,H],
    bias_ih_l0 [4H], bias_hh_l0 [4H] (gate order i|f|g|o), created in the same order and with the
    same U(-1/sqrt(H), 1/sqrt(H)) draws as torch.nn.LSTM, so a seed gives the same weights,
    reference checkpoints load unchanged and notebook code such as
    `model.lstm.bias_hh_l0.data[H:2*H] = 3` (LSTM_CAMELS_GB.ipynb:278) keeps working.
    """

    def __init__(self, input_size: int, hidden_size: int, batch_first: bool = True, num_layers: int = 1,
                 precision: str = "auto"):
        super().__init__()
        if num_layers != 1:
            raise NotImplementedError("hy2dl_b200.LSTM implements no_of_layers = 1 (every reference experiment uses 1)")
        if not batch_first:
            raise NotImplementedError("hy2dl_b200.LSTM is batch_first only, as the reference uses it")
        self.input_size, self.hidden_size, self.num_layers, self.batch_first = input_size, hidden_size, 1, True
        # "auto": the fastest kernels that meet the fp32 tolerances (rtol 1e-4) for this shape -- the tcgen05 path with
        # resident weights ("tf32x3": H = 64, input_size <= 31); otherwise "fp32", which is the streamed-weights tcgen05
        # chain for H = 128 / 256 and the CUDA-core kernels for H = 64 with wider inputs
        if precision == "auto":
            precision = "tf32x3" if (hidden_size == 64 and input_size <= 31) else "fp32"
        if precision not in ops.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(ops.PRECISIONS)} or 'auto', got '{precision}'")
        self.precision = precision
        H = hidden_size
        self.weight_ih_l0 = nn.Parameter(torch.empty(4 * H, input_size))
        self.weight_hh_l0 = nn.Parameter(torch.empty(4 * H, H))
        self.bias_ih_l0 = nn.Parameter(torch.empty(4 * H))
        self.bias_hh_l0 = nn.Parameter(torch.empty(4 * H))
        self.reset_parameters()

    def reset_parameters(self) -> None:
        stdv = 1.0 / math.sqrt(self.hidden_size) if self.hidden_size > 0 else 0
        for w in self.parameters():
            nn.init.uniform_(w, -stdv, stdv)

    def weights(self):
        return self.weight_ih_l0, self.weight_hh_l0, self.bias_ih_l0, self.bias_hh_l0

    @property
    def layout(self) -> str:
        """Layout of the native streams this module reads and writes ("rowmajor" | "ri32")."""
        return ops.layout_of(self.precision)

    def run_native(self, xT: torch.Tensor, dh_t0: int = 0, want_seq: bool = True, B: Optional[int] = None):
        """xT: native input ([T,B,IP], or RI32 for precision tf32x3 with B given) ->
        (h_seq in the same layout or empty, h_last [B,H])."""
        return ops.lstm(xT, *self.weights(), self.input_size, dh_t0, want_seq, self.precision, B)

    def forward(self, x: torch.Tensor, hx: Optional[Tuple[torch.Tensor, torch.Tensor]] = None):
        """nn.LSTM-compatible call: x [B,T,I] -> (out [B,T,H], (h_n [1,B,H], c_n=None)).

        The initial state is zero, which is the only way the reference ever calls it
        (cudalstm.py:43-50); `hx` is accepted for signature compatibility and not read.
        """
        if x.dim() != 3 or x.shape[2] != self.input_size:
            raise ValueError(f"expected x of shape [B, T, {self.input_size}], got {tuple(x.shape)}")
        B = x.shape[0]
        h_seq, h_last = self.run_native(ops.pack_x(x, self.layout), B=B)
        if ops.is_ri32(h_seq):
            h_seq = ops.from_ri32(h_seq, B)
        return h_seq.permute(1, 0, 2), (h_last.unsqueeze(0), None)

    def extra_repr(self) -> str:
        return f"{self.input_size}, {self.hidden_size}, batch_first=True, precision={self.precision}"
