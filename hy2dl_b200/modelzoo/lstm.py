"""Single-layer LSTM module with torch.nn.LSTM's parameter names, shapes and initialisation,
executed by the hy2dl_b200 recurrent kernels instead of cuDNN / oneDNN."""
from __future__ import annotations

import math
import os
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
        # The reference takes any sizes (cudalstm.py:18-24).  The kernels are built for 64 / 128 / 256 hidden units:
        # other sizes run zero-padded to the next of those (exact: a padded unit has z = 0 for every gate, hence
        # i = f = o = 1/2, g = 0, c = h = 0 for all t, and feeds zero columns of W_hh), see kernel_weights().
        if not (1 <= hidden_size <= 256):
            raise ValueError(f"hy2dl_b200.LSTM: hidden_size must be in [1, 256] (got {hidden_size})")
        if not (1 <= input_size <= 64):
            raise ValueError(f"hy2dl_b200.LSTM: input_size must be in [1, 64] (got {input_size})")
        self.kernel_hidden = 64 if hidden_size <= 64 else (128 if hidden_size <= 128 else 256)
        # hidden_size <= 64 with 32 .. 64 inputs: the resident-weights kernels take input_size <= 31, and the CUDA-core
        # kernels they used to fall back to are slower than the tensor-core chain run zero-padded to 128 units
        # (I = 41, T = 365, forward + BPTT + weight gradients: 40.7 vs 33.7 ms at a batch of 18 944, 5.9 vs 4.1 ms at 256).
        # HY2DL_FP32_TC=0 (CUDA-core kernels everywhere) keeps 64.
        if self.kernel_hidden == 64 and input_size > 31 and precision != "tf32x3" and os.environ.get("HY2DL_FP32_TC") != "0":
            self.kernel_hidden = 128
        # "auto": the fastest kernels that meet the fp32 tolerances (rtol 1e-4) for this shape -- the tcgen05 path with
        # resident weights ("tf32x3": H <= 64, input_size <= 31); otherwise "fp32", which is the tcgen05 chain of the
        # H = 128 / 256 kernels (streamed weights, or unit split at small batches).  "tf32" is the throughput mode
        # (single-pass TF32 operands on the same tensor-core kernels; looser, stated tolerance -- never picked by "auto")
        auto = precision == "auto"
        if auto:
            precision = "tf32x3" if (self.kernel_hidden == 64 and input_size <= 31) else "fp32"
        if precision not in ops.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(ops.PRECISIONS)} or 'auto', got '{precision}'")
        if precision == "tf32x3" and not (self.kernel_hidden == 64 and input_size <= 31):
            raise ValueError("hy2dl_b200.LSTM: precision 'tf32x3' (weights resident on chip) takes hidden_size <= 64 and "
                             f"input_size <= 31 (got {hidden_size}, {input_size}); use 'auto' or 'fp32'")
        if precision == "tf32" and self.kernel_hidden == 64 and input_size > 31:       # only with HY2DL_FP32_TC=0
            raise ValueError("hy2dl_b200.LSTM: precision 'tf32' (throughput mode) runs on the tensor-core kernels only "
                             f"(got hidden_size {hidden_size}, input_size {input_size} with HY2DL_FP32_TC=0)")
        # Small batches of a default ("auto") H <= 64 model run on the unit-split kernels zero-padded to 128 units: a 128-sequence
        # tile of the resident-weights kernels occupies ONE SM whatever the batch (the reference's batch of 256 = 2 of 148 SMs),
        # the unit-split kernels spread a tile over 8 SMs.  Forward + BPTT + weight gradients, T = 365, I = 25: 3.9 vs 5.5 ms at a
        # batch of 256, 4.4 vs 5.7 at 1 024, 5.2 vs 5.8 at 2 304 (one round of 18 groups).  Chosen per call from the batch size
        # (layout_for); both families meet the same fp32 tolerances.  HY2DL_SMALL_BATCH_US=0 switches it off.
        self._small_family = (auto and precision == "tf32x3" and os.environ.get("HY2DL_SMALL_BATCH_US") != "0"
                              and os.environ.get("HY2DL_FP32_TC") != "0" and os.environ.get("HY2DL_RM_US") != "0")
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

    SMALL_BATCH_MAX = 4736      # one round of the 37 unit-split groups of a 148-SM GPU at 64 units

    def family_for(self, B: int):
        """(kernel hidden size, precision, layout) of the kernels a batch of B sequences runs on."""
        if self._small_family and B <= self.SMALL_BATCH_MAX:
            return 64, "fp32", "rowmajor"
        return self.kernel_hidden, self.precision, self.layout

    def layout_for(self, B: int) -> str:
        """Layout the native input of a batch of B sequences must be packed in (the sampler / pack_x write it)."""
        return self.family_for(B)[2]

    def kernel_weights(self, HK: Optional[int] = None):
        """The four parameters at the kernels' hidden size HK (default kernel_hidden): unit u of gate g sits at row
        g*HK + u, padded rows / columns are zero.  Differentiable torch indexing on the (small) weights, so the
        gradients of the padded tensors flow back to the real parameters."""
        H, HK = self.hidden_size, HK or self.kernel_hidden
        if H == HK:
            return self.weights()
        w_ih, w_hh, b_ih, b_hh = self.weights()
        pad_rows = lambda w: torch.nn.functional.pad(w.view(4, H, -1), (0, 0, 0, HK - H)).reshape(4 * HK, -1)
        w_ih_p = pad_rows(w_ih)
        w_hh_p = torch.nn.functional.pad(pad_rows(w_hh), (0, HK - H))
        return w_ih_p, w_hh_p, pad_rows(b_ih.view(-1, 1)).view(-1), pad_rows(b_hh.view(-1, 1)).view(-1)

    def pad_head(self, w_head: torch.Tensor, HK: Optional[int] = None) -> torch.Tensor:
        """A Linear(H, C) weight [C, H] at the kernels' hidden size [C, HK] (zero columns for the padded units)."""
        H, HK = self.hidden_size, HK or self.kernel_hidden
        return w_head if H == HK else torch.nn.functional.pad(w_head, (0, HK - H))

    @property
    def layout(self) -> str:
        """Layout of the native streams this module reads and writes ("rowmajor" | "ri32")."""
        return ops.layout_of(self.precision, self.kernel_hidden)

    def run_native(self, xT: torch.Tensor, dh_t0: int = 0, want_seq: bool = True, B: Optional[int] = None):
        """xT: native input ([T,B,IP], or RI32 for precision tf32x3 with B given) ->
        (h_seq in the same layout or empty, h_last [B,HK]).  The layout of xT says which family runs (layout_for)."""
        HK, precision = self.kernel_hidden, self.precision
        if self.layout == "ri32" and not ops.is_ri32(xT):            # row-major input to a resident-weights model: small-batch family
            if not self._small_family:
                raise ValueError("hy2dl_b200.LSTM: this model reads RI32 input (precision tf32x3); got a row-major tensor")
            precision = "fp32"
        h_seq, h_last = ops.lstm(xT, *self.kernel_weights(HK), self.input_size, dh_t0, want_seq, precision, B)
        return h_seq, h_last

    def forward(self, x: torch.Tensor, hx: Optional[Tuple[torch.Tensor, torch.Tensor]] = None):
        """nn.LSTM-compatible call: x [B,T,I] -> (out [B,T,H], (h_n [1,B,H], c_n=None)).

        The initial state is zero, which is the only way the reference ever calls it (cudalstm.py:43-50,
        hybrid.py:84-92: explicit zero tensors).  `hx` may therefore be None or all-zero tensors; anything else
        raises instead of being silently ignored.  c_n is not produced by the kernels (no reference caller reads it).
        """
        if x.dim() != 3 or x.shape[2] != self.input_size:
            raise ValueError(f"expected x of shape [B, T, {self.input_size}], got {tuple(x.shape)}")
        if hx is not None:
            if not (isinstance(hx, (tuple, list)) and len(hx) == 2):
                raise ValueError("hx must be a (h_0, c_0) pair")
            if any(bool(t.detach().ne(0).any()) for t in hx):
                raise NotImplementedError("hy2dl_b200.LSTM starts from h_0 = c_0 = 0 (as every reference model does); "
                                          "a non-zero initial state is not supported")
        B, H = x.shape[0], self.hidden_size
        h_seq, h_last = self.run_native(ops.pack_x(x, self.layout_for(B)), B=B)
        if ops.is_ri32(h_seq):
            h_seq = ops.from_ri32(h_seq, B)
        return h_seq[:, :, :H].permute(1, 0, 2), (h_last[:, :H].unsqueeze(0), None)

    def extra_repr(self) -> str:
        return f"{self.input_size}, {self.hidden_size}, batch_first=True, precision={self.precision}"
