"""torch.autograd.Function wrappers around the C-ABI kernels.

Tensors that travel between these functions are in the library's native time-major layouts
(include/hy2dl_b200.h); the model classes in hy2dl_b200.modelzoo turn them into the reference's
[batch, time, feature] views at the API boundary.  PyTorch is used for memory, streams and the
autograd graph only -- every arithmetic step on the path is one of our kernels.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import ParMap, call, ptr, require_cuda, stream

# "fp32" and "tf32x3" are the fp32-class parity paths; "tf32" is the throughput mode (the same tensor-core kernels without
# the correction terms of the 3-term TF32 split: 10-bit operands, fp32 accumulate; include/hy2dl_b200.h)
PRECISIONS = {"fp32": _lib.PREC_FP32, "tf32x3": _lib.PREC_TF32X3, "tf32": _lib.PREC_TF32}


def round_up(a: int, b: int) -> int:
    return (a + b - 1) // b * b


def _ri32_unit_perm(device) -> torch.Tensor:
    """[32 rows, 4 units] index: unit u of row r is stored at position u ^ (r & 3) (an involution)."""
    r = torch.arange(32, device=device).view(32, 1)
    u = torch.arange(4, device=device).view(1, 4)
    return u ^ (r & 3)


def to_ri32(x: torch.Tensor) -> torch.Tensor:
    """Row-major [T,B,F] (F % 32 == 0) -> RI32 [T,G,F/32,32,32] (G = ceil(B/32), zero padded rows, 8-float
    units of each row swizzled; include/hy2dl_b200.h). Torch ops; used at API boundaries and in tests,
    not on the training fast path (the sampler writes RI32 directly)."""
    T, B, F = x.shape
    assert F % 32 == 0, "RI32 streams are multiples of 32 features wide"
    G = (B + 31) // 32
    if G * 32 != B:
        x = torch.cat([x, x.new_zeros(T, G * 32 - B, F)], dim=1)
    y = x.view(T, G, 32, F // 32, 4, 8).permute(0, 1, 3, 2, 4, 5)                   # [T,G,FB,row,unit,8]
    idx = _ri32_unit_perm(x.device).view(1, 1, 1, 32, 4, 1).expand(T, G, F // 32, 32, 4, 8)
    return torch.gather(y, 4, idx).reshape(T, G, F // 32, 32, 32).contiguous()


def from_ri32(x: torch.Tensor, B: int) -> torch.Tensor:
    """RI32 [T,G,F/32,32,32] -> row-major [T,B,F]."""
    T, G, FB = x.shape[:3]
    y = x.reshape(T, G, FB, 32, 4, 8)
    y = torch.gather(y, 4, _ri32_unit_perm(x.device).view(1, 1, 1, 32, 4, 1).expand(T, G, FB, 32, 4, 8))
    return y.permute(0, 1, 3, 2, 4, 5).reshape(T, G * 32, FB * 32)[:, :B]


def _ws(kind: int, B: int, T: int, I: int, H: int, device, precision: int = 0) -> torch.Tensor:
    n = _lib.load().hy2dl_workspace_bytes(kind, B, T, I, H, precision)
    if n < 0:
        raise RuntimeError(_lib.last_error())
    return torch.empty(max(int(n), 16), dtype=torch.uint8, device=device)


# --------------------------------------------------------------------------------------------------
# layout packing (no gradient w.r.t. inputs: the reference never asks for dL/dx either)
# --------------------------------------------------------------------------------------------------
def is_ri32(t: torch.Tensor) -> bool:
    """Native LSTM-side streams are 3-D [T,B,F] (row-major) or 5-D [T,G,F/32,32,32] (RI32)."""
    return t.dim() == 5


def layout_of(precision: str, hidden: int = 64) -> str:
    """The stream layout a recurrent-kernel precision works in at the kernels' hidden size (64 / 128 / 256)."""
    return "ri32" if (precision == "tf32x3" or (precision == "tf32" and hidden == 64)) else "rowmajor"


def _ri32_family(precision: int, H: int) -> bool:
    return precision == _lib.PREC_TF32X3 or (precision == _lib.PREC_TF32 and H == 64)


def pack_x(x: torch.Tensor, layout: str = "rowmajor") -> torch.Tensor:
    """[B,T,I] -> native [T,B,IP] (layout "rowmajor") or RI32 [T,G,1,32,32] (layout "ri32", I <= 31)."""
    require_cuda(x, "x_lstm")
    x = x.detach().contiguous()
    B, T, I = x.shape
    IP = round_up(I, 8)
    if layout == "ri32":
        if I > 31:
            raise RuntimeError(f"hy2dl_b200: the tf32x3 path takes input_size <= 31 (got {I})")
        xT = torch.empty(T, (B + 31) // 32, 1, 32, 32, dtype=torch.float32, device=x.device)
        call("hy2dl_pack_x", ptr(x), B, T, I, 32, ptr(xT), _lib.LAYOUT_RI32, stream())
    else:
        xT = torch.empty(T, B, IP, dtype=torch.float32, device=x.device)
        call("hy2dl_pack_x", ptr(x), B, T, I, IP, ptr(xT), _lib.LAYOUT_ROWMAJOR, stream())
    return xT


def pack_forcing(xc: torch.Tensor) -> torch.Tensor:
    """[B,T,2|3|4] -> native [T,3,B] (P, PET, Tmean; two columns: temperature plane zero)."""
    require_cuda(xc, "x_conceptual")
    xc = xc.detach().contiguous()
    B, T, nc = xc.shape
    forc = torch.empty(T, 3, B, dtype=torch.float32, device=xc.device)
    call("hy2dl_pack_forcing", ptr(xc), B, T, nc, ptr(forc), stream())
    return forc


# --------------------------------------------------------------------------------------------------
# LSTM
# --------------------------------------------------------------------------------------------------
def _consume_tape(ctx) -> None:
    """BPTT overwrites the gate tape with dG in place (it is 2/3 of the tape: a copy would cost 7 GB per step at the
    benchmark batch), so a graph can be back-propagated once; a second pass (retain_graph=True) must fail loudly
    instead of reading gradients where it expects activations."""
    if getattr(ctx, "tape_consumed", False):
        raise RuntimeError("hy2dl_b200: this LSTM graph was already back-propagated; its tape is overwritten in place "
                           "(retain_graph / double backward are not supported) -- run the forward pass again")
    ctx.tape_consumed = True


class LstmFunction(torch.autograd.Function):
    """h_seq and h_last [B,H] from xT; one layer, h0 = c0 = 0.

    Row-major precisions: xT [T,B,IP] -> h_seq [T,B,H].  tf32x3: xT RI32 [T,G,1,32,32] ->
    h_seq RI32 [T,G,H/32,32,32] (B = number of real rows, given explicitly).
    `dh_t0`: the caller promises that the gradient flowing into h_seq is zero (and may be
    uninitialised memory) for steps t < dh_t0 -- the hybrid head only feeds steps >= warm-up.
    """

    @staticmethod
    def forward(ctx, xT, w_ih, w_hh, b_ih, b_hh, I: int, dh_t0: int, want_seq: bool, precision: int, B: int):
        for n, t in (("xT", xT), ("weight_ih", w_ih), ("weight_hh", w_hh), ("bias_ih", b_ih), ("bias_hh", b_hh)):
            require_cuda(t, n)
        H = w_hh.shape[1]
        ri = _ri32_family(precision, H)
        if ri != is_ri32(xT):
            raise RuntimeError("hy2dl_b200: the resident-weights path (H = 64, precision 'tf32x3' / 'tf32') works on RI32 inputs and "
                               f"the others on row-major ones (got a {xT.dim()}-D xT); pack / gather with layout=ops.layout_of(precision, H)")
        dev = xT.device
        if ri:
            T, G, IP = xT.shape[0], xT.shape[1], xT.shape[2] * 32
            seq = lambda F: torch.empty(T, G, F // 32, 32, 32, dtype=torch.float32, device=dev)
        else:
            T, B, IP = xT.shape
            seq = lambda F: torch.empty(T, B, F, dtype=torch.float32, device=dev)
        need_grad = any(ctx.needs_input_grad[1:5])
        ctx.set_materialize_grads(False)     # an unused output (h_seq or h_last) arrives as None, not as a zero tensor
        w_ih, w_hh, b_ih, b_hh = (t.detach().contiguous() for t in (w_ih, w_hh, b_ih, b_hh))
        h_seq = seq(H) if (want_seq or need_grad) else None
        h_last = torch.empty(B, H, dtype=torch.float32, device=dev)
        # c_seq and gates are tapes only the kernels read; the H = 128 / 256 tensor-core chain stores them blocked by
        # 32 sequences (include/hy2dl_b200.h, "RB8"), so they hold whole groups of rows
        # (H = 64 at small batches as well: lstm_fp32.cu us64_chain; the CUDA-core kernels use the first T*B*F floats)
        tape = seq if ri else (lambda F: torch.empty(T, (B + 31) // 32 * 32, F, dtype=torch.float32, device=dev))
        c_seq = tape(H) if need_grad else None
        gates = tape(4 * H) if need_grad else None
        ws = _ws(_lib.WS_LSTM_FWD, B, T, I, H, dev, precision)
        call("hy2dl_lstm_fwd", ptr(xT), B, T, I, IP, H, ptr(w_ih), ptr(w_hh), ptr(b_ih), ptr(b_hh), ptr(h_seq),
             ptr(c_seq), ptr(gates), ptr(h_last), ptr(ws), ws.numel(), precision, None, stream())
        ctx.save_for_backward(xT, w_hh, h_seq, c_seq, gates)
        ctx.meta = (B, T, I, IP, H, dh_t0, precision, want_seq)
        if not want_seq:
            none = torch.empty(0, device=dev)
            ctx.mark_non_differentiable(none)
            return none, h_last
        return h_seq, h_last

    @staticmethod
    def backward(ctx, dh_seq, dh_last):
        xT, w_hh, h_seq, c_seq, gates = ctx.saved_tensors
        B, T, I, IP, H, dh_t0, precision, want_seq = ctx.meta
        dev = xT.device
        if not want_seq or dh_seq is None or dh_seq.numel() == 0:
            dh_seq = None
        else:
            dh_seq = dh_seq.contiguous()
        if dh_last is not None:
            dh_last = dh_last.contiguous()
        if dh_seq is None and dh_last is None:
            return (None,) * 10
        _consume_tape(ctx)
        ws_b = _ws(_lib.WS_LSTM_BWD, B, T, I, H, dev, precision)
        # dG overwrites the gate tape in place (the kernel reads each element before writing it)
        dG = gates
        call("hy2dl_lstm_bwd", B, T, H, ptr(w_hh), ptr(gates), ptr(c_seq), ptr(dh_seq), dh_t0, ptr(dh_last), ptr(dG),
             ptr(ws_b), ws_b.numel(), precision, None, stream())
        dw_ih = torch.empty(4 * H, I, dtype=torch.float32, device=dev)
        dw_hh = torch.empty(4 * H, H, dtype=torch.float32, device=dev)
        db = torch.empty(4 * H, dtype=torch.float32, device=dev)
        ws_w = _ws(_lib.WS_LSTM_WGRAD, B, T, I, H, dev, precision)
        call("hy2dl_lstm_wgrad", ptr(dG), ptr(h_seq), ptr(xT), B, T, I, IP, H, ptr(dw_ih), ptr(dw_hh), ptr(db),
             ptr(ws_w), ws_w.numel(), precision, None, stream())
        return None, dw_ih, dw_hh, db, db.clone(), None, None, None, None, None


def lstm(xT, w_ih, w_hh, b_ih, b_hh, input_size: int, dh_t0: int = 0, want_seq: bool = True,
         precision: str = "fp32", B: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """B: number of sequences; required for RI32 inputs (whose row count is padded to 32)."""
    if B is None:
        if is_ri32(xT):
            raise ValueError("hy2dl_b200.ops.lstm: pass B (the number of sequences) with an RI32 input")
        B = xT.shape[1]
    return LstmFunction.apply(xT, w_ih, w_hh, b_ih, b_hh, input_size, dh_t0, want_seq, PRECISIONS[precision], B)


# --------------------------------------------------------------------------------------------------
# head + parameter mapping
# --------------------------------------------------------------------------------------------------
@dataclass
class ParLayout:
    """Where each conceptual / routing parameter's plane lives inside the flat par_sim buffer."""
    pm: ParMap
    names: List[str]
    routing_names: List[str]
    n_floats: int
    B: int
    T: int

    @property
    def T_sim(self) -> int:
        return self.T - self.pm.warmup

    @property
    def N(self) -> int:
        return self.B * self.pm.n_members

    def plane(self, flat: torch.Tensor, k: int) -> torch.Tensor:
        """View of parameter k (routing follows the conceptual ones): [T_sim,N] dynamic, [N] static, [B] routing."""
        off = self.pm.sim_off[k]
        if k >= self.pm.n_par:
            return flat[off:off + self.B]
        if self.pm.dynamic[k]:
            return flat[off:off + self.T_sim * self.N].view(self.T_sim, self.N)
        return flat[off:off + self.N]


def make_layout(ranges: Dict[str, Tuple[float, float]], dynamic: Sequence[str], n_members: int, warmup: int,
                routing_ranges: Optional[Dict[str, Tuple[float, float]]], B: int, T: int) -> ParLayout:
    pm = ParMap()
    names = list(ranges)
    if len(names) > _lib.MAX_PAR:
        raise ValueError(f"at most {_lib.MAX_PAR} conceptual parameters are supported")
    pm.n_par, pm.n_members, pm.warmup = len(names), n_members, warmup
    rnames = list(routing_ranges) if routing_ranges else []
    pm.n_routing = len(rnames)
    for k, n in enumerate(names):
        pm.dynamic[k] = 1 if n in dynamic else 0
        pm.lo[k], pm.hi[k] = ranges[n]
    for r, n in enumerate(rnames):
        pm.lo[len(names) + r], pm.hi[len(names) + r] = routing_ranges[n]
    nf = _lib.load().hy2dl_parmap_layout(C.byref(pm), B, T)
    if nf < 0:
        raise RuntimeError("hy2dl_parmap_layout failed")
    return ParLayout(pm, names, rnames, int(nf), B, T)


class HeadMapFunction(torch.autograd.Function):
    """(h_seq, W, b) -> (par_sim flat planes, par_warm [P,N]); par_warm carries no gradient."""

    @staticmethod
    def forward(ctx, h_seq, w, bias, layout: ParLayout):
        require_cuda(h_seq, "h_seq"); require_cuda(w, "linear.weight"); require_cuda(bias, "linear.bias")
        ri = is_ri32(h_seq)
        T, B, H = h_seq.shape[0], layout.B, w.shape[1]
        lay = _lib.LAYOUT_RI32 if ri else _lib.LAYOUT_ROWMAJOR
        dev = h_seq.device
        h_seq = h_seq.contiguous()
        w_c, b_c = w.detach().contiguous(), bias.detach().contiguous()
        par_sim = torch.empty(layout.n_floats, dtype=torch.float32, device=dev)
        par_warm = torch.empty(layout.pm.n_par, layout.N, dtype=torch.float32, device=dev)
        call("hy2dl_head_map_fwd", ptr(h_seq), B, T, H, ptr(w_c), ptr(b_c), C.byref(layout.pm), ptr(par_sim),
             ptr(par_warm), lay, stream())
        ctx.save_for_backward(h_seq, w_c, par_sim)
        ctx.layout = layout
        ctx.mark_non_differentiable(par_warm)
        ctx.set_materialize_grads(False)
        return par_sim, par_warm

    @staticmethod
    def backward(ctx, dpar_sim, _dpar_warm):
        if dpar_sim is None:
            return None, None, None, None
        h_seq, w, par_sim = ctx.saved_tensors
        layout = ctx.layout
        T, B, H = h_seq.shape[0], layout.B, w.shape[1]
        lay = _lib.LAYOUT_RI32 if is_ri32(h_seq) else _lib.LAYOUT_ROWMAJOR
        dev = h_seq.device
        dpar_sim = dpar_sim.contiguous()
        dh_seq = torch.empty_like(h_seq)  # rows < warmup stay uninitialised (dh_t0)
        dw = torch.empty_like(w)
        db = torch.empty(w.shape[0], dtype=torch.float32, device=dev)
        ws = torch.empty(int(_lib.load().hy2dl_head_bwd_workspace_bytes(H)), dtype=torch.uint8, device=dev)
        call("hy2dl_head_map_bwd", ptr(h_seq), ptr(par_sim), ptr(dpar_sim), B, T, H, ptr(w), C.byref(layout.pm),
             ptr(dh_seq), ptr(dw), ptr(db), ptr(ws), ws.numel(), lay, stream())
        return dh_seq, dw, db, None


def head_map(h_seq, w, bias, layout: ParLayout):
    return HeadMapFunction.apply(h_seq, w, bias, layout)


FUSED_HEAD_MAX_COLUMNS = 32


class LstmHeadFunction(torch.autograd.Function):
    """LSTM + head in one forward kernel (precision tf32x3, <= 32 head columns): RI32 xT ->
    (par_sim flat planes, par_warm [P,N]).  The hidden sequence never leaves the kernels' own layout."""

    @staticmethod
    def forward(ctx, xT, w_ih, w_hh, b_ih, b_hh, w_head, b_head, I: int, layout: ParLayout, precision: int = _lib.PREC_TF32X3):
        for n, t in (("xT", xT), ("weight_ih", w_ih), ("weight_hh", w_hh), ("bias_ih", b_ih), ("bias_hh", b_hh),
                     ("linear.weight", w_head), ("linear.bias", b_head)):
            require_cuda(t, n)
        if not is_ri32(xT):
            raise RuntimeError("hy2dl_b200: the fused LSTM + head path works on RI32 inputs (H = 64, precision 'tf32x3' / 'tf32')")
        B, H, dev = layout.B, w_hh.shape[1], xT.device
        T, G, IP = xT.shape[0], xT.shape[1], xT.shape[2] * 32
        seq = lambda F: torch.empty(T, G, F // 32, 32, 32, dtype=torch.float32, device=dev)
        need_grad = any(ctx.needs_input_grad[1:7])
        w_ih, w_hh, b_ih, b_hh, w_head, b_head = (t.detach().contiguous() for t in (w_ih, w_hh, b_ih, b_hh, w_head, b_head))
        h_seq = seq(H) if need_grad else None
        c_seq = seq(H) if need_grad else None
        gates = seq(4 * H) if need_grad else None
        par_sim = torch.empty(layout.n_floats, dtype=torch.float32, device=dev)
        par_warm = torch.empty(layout.pm.n_par, layout.N, dtype=torch.float32, device=dev)
        head = _lib.HeadFwd(ptr(w_head), ptr(b_head), C.pointer(layout.pm), ptr(par_sim), ptr(par_warm))
        ws = _ws(_lib.WS_LSTM_FWD, B, T, I, H, dev, precision)
        call("hy2dl_lstm_fwd", ptr(xT), B, T, I, IP, H, ptr(w_ih), ptr(w_hh), ptr(b_ih), ptr(b_hh), ptr(h_seq),
             ptr(c_seq), ptr(gates), None, ptr(ws), ws.numel(), precision, C.byref(head), stream())
        ctx.save_for_backward(xT, w_hh, w_head, h_seq, c_seq, gates, par_sim)
        ctx.meta = (B, T, I, IP, H, precision)
        ctx.layout = layout
        ctx.mark_non_differentiable(par_warm)
        ctx.set_materialize_grads(False)
        return par_sim, par_warm

    @staticmethod
    def backward(ctx, dpar_sim, _dpar_warm):
        if dpar_sim is None:
            return (None,) * 10
        xT, w_hh, w_head, h_seq, c_seq, gates, par_sim = ctx.saved_tensors
        B, T, I, IP, H, precision = ctx.meta
        layout, dev = ctx.layout, xT.device
        dpar_sim = dpar_sim.contiguous()
        _consume_tape(ctx)
        # BPTT with the head adjoint dz * W_head accumulated by the kernel's own MMAs (no dh_seq exists); the kernel
        # also leaves dz (RI32) for the weight-gradient GEMM
        G = xT.shape[1]
        dz = torch.empty(T, G, 1, 32, 32, dtype=torch.float32, device=dev)
        dw_head = torch.empty_like(w_head)
        db_head = torch.empty(w_head.shape[0], dtype=torch.float32, device=dev)
        hb = _lib.HeadBwd(ptr(w_head), C.pointer(layout.pm), ptr(par_sim), ptr(dpar_sim), ptr(dz))
        ws_b = _ws(_lib.WS_LSTM_BWD, B, T, I, H, dev, precision)
        dG = gates   # in place
        call("hy2dl_lstm_bwd", B, T, H, ptr(w_hh), ptr(gates), ptr(c_seq), None, layout.pm.warmup, None, ptr(dG),
             ptr(ws_b), ws_b.numel(), precision, C.byref(hb), stream())
        # all weight gradients in one GEMM over the tape: dW_ih, dW_hh, db, dW_head and db_head
        dw_ih = torch.empty(4 * H, I, dtype=torch.float32, device=dev)
        dw_hh = torch.empty(4 * H, H, dtype=torch.float32, device=dev)
        db = torch.empty(4 * H, dtype=torch.float32, device=dev)
        hw = _lib.HeadWgrad(ptr(dz), layout.pm.warmup, w_head.shape[0], ptr(dw_head), ptr(db_head))
        ws_w = _ws(_lib.WS_LSTM_WGRAD, B, T, I, H, dev, precision)
        call("hy2dl_lstm_wgrad", ptr(dG), ptr(h_seq), ptr(xT), B, T, I, IP, H, ptr(dw_ih), ptr(dw_hh), ptr(db),
             ptr(ws_w), ws_w.numel(), precision, C.byref(hw), stream())
        return None, dw_ih, dw_hh, db, db.clone(), dw_head, db_head, None, None, None


def lstm_head(xT, w_ih, w_hh, b_ih, b_hh, w_head, b_head, input_size: int, layout: ParLayout, precision: str = "tf32x3"):
    return LstmHeadFunction.apply(xT, w_ih, w_hh, b_ih, b_hh, w_head, b_head, input_size, layout, PRECISIONS[precision])


# --------------------------------------------------------------------------------------------------
# conceptual models
# --------------------------------------------------------------------------------------------------
def _ptr_array(ptrs: Sequence[Optional[int]]):
    arr = (C.c_void_p * _lib.MAX_PAR)()
    for k, p in enumerate(ptrs):
        arr[k] = p
    return arr


def _i64_array(vals: Sequence[int]):
    arr = (C.c_int64 * _lib.MAX_PAR)()
    for k, v in enumerate(vals):
        arr[k] = v
    return arr


class ConceptualFunction(torch.autograd.Function):
    """Bucket model over forcing rows [t0, t0+Tn).

    par_flat: flat buffer, parameter k's plane at offs[k] (None: parameter absent), dynamic[k] says
    whether the plane is [Tn,N] or [N].  Returns (states [Tn,5,N], q [Tn,B], routing).

    `routing_span` = (offset, length) of the routing parameters inside par_flat, or None: they are handed on as a
    small copy (third output) and their gradient is written back into this function's own gradient of par_flat.
    Slicing the big buffer in autograd instead would cost a zero-filled full-size gradient and a full-size add per
    slice (2 x 110 MB fills and 2 x 330 MB adds per step at the benchmark batch).
    """

    @staticmethod
    def forward(ctx, forc, par_flat, init, model: int, B: int, m: int, t0: int, Tn: int, offs, dynamic,
                routing_span=None):
        require_cuda(forc, "forcing"); require_cuda(par_flat, "parameters")
        dev = forc.device
        N = B * m
        forc = forc.contiguous()
        par_flat = par_flat.contiguous()
        if init is not None:
            init = require_cuda(init, "initial_states").contiguous()
        base = par_flat.data_ptr()
        ptrs = [None if o is None else base + 4 * o for o in offs]
        ts = [0 if (o is None or not d) else N for o, d in zip(offs, dynamic)]
        states = torch.empty(Tn, _lib.N_STATES, N, dtype=torch.float32, device=dev)
        q = torch.empty(Tn, B, dtype=torch.float32, device=dev)
        call("hy2dl_conceptual_fwd", model, ptr(forc), B, t0, Tn, m, _ptr_array(ptrs), _i64_array(ts), ptr(init),
             ptr(states), ptr(q), stream())
        ctx.save_for_backward(forc, par_flat, init, states)
        # the hybrid never differentiates the storages: without this, autograd hands the adjoint kernel a zero-filled
        # [Tn,5,N] tensor (17 M floats at the benchmark batch: one fill launch plus five extra plane reads per step)
        ctx.set_materialize_grads(False)
        ctx.meta = (model, B, m, t0, Tn, tuple(offs), tuple(dynamic), routing_span)
        routing = None
        if routing_span is not None:
            routing = par_flat.detach()[routing_span[0]:routing_span[0] + routing_span[1]].clone()
        return states, q, routing

    @staticmethod
    def backward(ctx, dstates, dq, drouting):
        forc, par_flat, init, states = ctx.saved_tensors
        model, B, m, t0, Tn, offs, dynamic, routing_span = ctx.meta
        if dstates is None and dq is None and drouting is None:
            return (None,) * 11
        dev = forc.device
        N = B * m
        base = par_flat.data_ptr()
        ptrs = [None if o is None else base + 4 * o for o in offs]
        ts = [0 if (o is None or not d) else N for o, d in zip(offs, dynamic)]
        dpar = torch.zeros_like(par_flat)  # zeros: the padding between planes must not leak garbage into autograd
        dbase = dpar.data_ptr()
        dptrs = [None if o is None else dbase + 4 * o for o in offs]
        dq = torch.zeros(Tn, B, dtype=torch.float32, device=dev) if dq is None else dq.contiguous()
        dstates = None if dstates is None else dstates.contiguous()
        dinit = torch.zeros_like(init) if (init is not None and ctx.needs_input_grad[2]) else None
        call("hy2dl_conceptual_bwd", model, ptr(forc), B, t0, Tn, m, _ptr_array(ptrs), _i64_array(ts), ptr(init),
             ptr(states), ptr(dq), ptr(dstates), _ptr_array(dptrs), ptr(dinit), stream())
        if drouting is not None:
            dpar[routing_span[0]:routing_span[0] + routing_span[1]] = drouting
        return None, dpar, dinit, None, None, None, None, None, None, None, None


def conceptual(forc, par_flat, init, model: int, B: int, m: int, t0: int, Tn: int, offs, dynamic):
    states, q, _ = ConceptualFunction.apply(forc, par_flat, init, model, B, m, t0, Tn, offs, dynamic, None)
    return states, q


def conceptual_routing(forc, par_flat, init, model: int, B: int, m: int, t0: int, Tn: int, offs, dynamic, routing_span):
    """As `conceptual`, additionally returning the routing-parameter span of par_flat as its own small tensor."""
    return ConceptualFunction.apply(forc, par_flat, init, model, B, m, t0, Tn, offs, dynamic, routing_span)


# --------------------------------------------------------------------------------------------------
# routing
# --------------------------------------------------------------------------------------------------
class UHFunction(torch.autograd.Function):
    """q [T,B], alpha [B], beta [B] -> routed y [T,B]."""

    @staticmethod
    def forward(ctx, q, alpha, beta):
        require_cuda(q, "discharge"); require_cuda(alpha, "alpha"); require_cuda(beta, "beta")
        T, B = q.shape
        q, alpha, beta = q.contiguous(), alpha.contiguous(), beta.contiguous()
        uh = torch.empty(15, B, dtype=torch.float32, device=q.device)
        y = torch.empty_like(q)
        call("hy2dl_uh_fwd", ptr(q), ptr(alpha), ptr(beta), B, T, ptr(uh), ptr(y), stream())
        ctx.save_for_backward(q, alpha, beta, uh)
        return y

    @staticmethod
    def backward(ctx, dy):
        q, alpha, beta, uh = ctx.saved_tensors
        T, B = q.shape
        dy = dy.contiguous()
        dq = torch.empty_like(q)
        da = torch.empty_like(alpha)
        db = torch.empty_like(beta)
        call("hy2dl_uh_bwd", ptr(q), ptr(alpha), ptr(beta), ptr(uh), ptr(dy), B, T, ptr(dq), ptr(da), ptr(db), stream())
        return dq, da, db


def uh_route(q, alpha, beta):
    return UHFunction.apply(q, alpha, beta)


class UHPackedFunction(torch.autograd.Function):
    """q [T,B] and one tensor holding alpha at [0,B) and beta at [beta_off, beta_off+B) -> routed y [T,B]; the two
    gradients are written straight into one gradient tensor of the same shape (no slicing in autograd)."""

    @staticmethod
    def forward(ctx, q, ab, beta_off: int):
        require_cuda(q, "discharge"); require_cuda(ab, "routing parameters")
        T, B = q.shape
        q, ab = q.contiguous(), ab.contiguous()
        uh = torch.empty(15, B, dtype=torch.float32, device=q.device)
        y = torch.empty_like(q)
        call("hy2dl_uh_fwd", ptr(q), ptr(ab), ab.data_ptr() + 4 * beta_off, B, T, ptr(uh), ptr(y), stream())
        ctx.save_for_backward(q, ab, uh)
        ctx.beta_off = beta_off
        return y

    @staticmethod
    def backward(ctx, dy):
        q, ab, uh = ctx.saved_tensors
        T, B = q.shape
        off = ctx.beta_off
        dy = dy.contiguous()
        dq = torch.empty_like(q)
        dab = torch.zeros_like(ab) if ab.numel() != 2 * B else torch.empty_like(ab)   # padding between the planes stays zero
        call("hy2dl_uh_bwd", ptr(q), ptr(ab), ab.data_ptr() + 4 * off, ptr(uh), ptr(dy), B, T, ptr(dq), ptr(dab),
             dab.data_ptr() + 4 * off, stream())
        return dq, dab, None


def uh_route_packed(q, ab, beta_off: int):
    return UHPackedFunction.apply(q, ab, beta_off)


# --------------------------------------------------------------------------------------------------
# losses
# --------------------------------------------------------------------------------------------------
def _allreduce_sum(t: torch.Tensor, group):
    """Sum the masked partial sums over the data-parallel group so that every rank forms the
    GLOBAL mean. group=None: single process, nothing to do."""
    if group is None:
        return
    import torch.distributed as dist
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=None if group == "world" else group)


def _loss_acc(device) -> torch.Tensor:
    """Zeroed accumulator of the loss reductions: sums at [0,4), then ticket and per-block partials
    (include/hy2dl_b200.h)."""
    return torch.zeros(int(_lib.load().hy2dl_loss_acc_floats()), dtype=torch.float32, device=device)


class NseLossFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y_sim, y_obs, std, group):
        require_cuda(y_sim, "y_sim"); require_cuda(y_obs, "y_obs"); require_cuda(std, "per_basin_target_std")
        n = y_sim.numel()
        acc = _loss_acc(y_sim.device)
        call("hy2dl_loss_nse_reduce", ptr(y_sim), ptr(y_obs), ptr(std), n, ptr(acc), stream())
        _allreduce_sum(acc[:4], group)
        loss = torch.empty((), dtype=torch.float32, device=y_sim.device)
        call("hy2dl_loss_nse_finish", ptr(y_sim), ptr(y_obs), ptr(std), n, ptr(acc), None, ptr(loss), None, stream())
        ctx.save_for_backward(y_sim, y_obs, std, acc)
        return loss

    @staticmethod
    def backward(ctx, g):
        y_sim, y_obs, std, acc = ctx.saved_tensors
        dy = torch.empty_like(y_sim)
        tmp = torch.empty((), dtype=torch.float32, device=y_sim.device)
        g = g.contiguous().to(torch.float32)
        call("hy2dl_loss_nse_finish", ptr(y_sim), ptr(y_obs), ptr(std), y_sim.numel(), ptr(acc), ptr(g), ptr(tmp),
             ptr(dy), stream())
        return dy, None, None, None


class WrmseLossFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y_sim, y_obs, group):
        require_cuda(y_sim, "y_sim"); require_cuda(y_obs, "y_obs")
        n = y_sim.numel()
        acc = _loss_acc(y_sim.device)
        call("hy2dl_loss_wrmse_reduce", ptr(y_sim), ptr(y_obs), n, ptr(acc), stream())
        _allreduce_sum(acc[:4], group)
        loss = torch.empty((), dtype=torch.float32, device=y_sim.device)
        call("hy2dl_loss_wrmse_finish", ptr(y_sim), ptr(y_obs), n, ptr(acc), None, ptr(loss), None, stream())
        ctx.save_for_backward(y_sim, y_obs, acc)
        return loss

    @staticmethod
    def backward(ctx, g):
        y_sim, y_obs, acc = ctx.saved_tensors
        dy = torch.empty_like(y_sim)
        tmp = torch.empty((), dtype=torch.float32, device=y_sim.device)
        g = g.contiguous().to(torch.float32)
        call("hy2dl_loss_wrmse_finish", ptr(y_sim), ptr(y_obs), y_sim.numel(), ptr(acc), ptr(g), ptr(tmp), ptr(dy),
             stream())
        return dy, None, None
