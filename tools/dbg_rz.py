"""Is the tcgen05 fp32 accumulation round-to-nearest or truncating?  Long positive accumulation chains in
the wgrad kernel: a truncating accumulator shows a systematic negative relative error growing with length."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hy2dl_b200 import _lib, ops
dev = torch.device("cuda:0")
H, I, IP = 64, 32, 32
for (B, T) in [(4736, 4), (18944, 16), (18944, 64)]:
    g = torch.Generator(device=dev); g.manual_seed(1)
    if os.environ.get("DBG_SIGNED"):
        dG = torch.randn(T, B, 4 * H, device=dev, generator=g); h = torch.randn(T, B, H, device=dev, generator=g)
        x = torch.randn(T, B, IP, device=dev, generator=g)
    else:
        dG = torch.rand(T, B, 4 * H, device=dev, generator=g) + 0.5
        h = torch.rand(T, B, H, device=dev, generator=g) + 0.5
        x = torch.rand(T, B, IP, device=dev, generator=g) + 0.5
    dwi = torch.empty(4 * H, I, device=dev); dwh = torch.empty(4 * H, H, device=dev); dbb = torch.empty(4 * H, device=dev)
    ws = ops._ws(_lib.WS_LSTM_WGRAD, B, T, I, H, dev, _lib.PREC_TF32X3)
    args = [ops.to_ri32(a) for a in (dG, h, x)]
    _lib.call("hy2dl_lstm_wgrad", *[a.data_ptr() for a in args], B, T, I, IP, H, dwi.data_ptr(), dwh.data_ptr(),
              dbb.data_ptr(), ws.data_ptr(), ws.numel(), _lib.PREC_TF32X3, None, _lib.stream())
    torch.cuda.synchronize()
    d = dG.double().reshape(T, B, H, 4).permute(0, 1, 3, 2).reshape(T * B, 4 * H)
    ref_i = d.t() @ x.double().reshape(T * B, IP)
    ref_b = d.sum(0)
    rel = ((dwi.double() - ref_i) / ref_i.abs().max()) if os.environ.get("DBG_SIGNED") else ((dwi.double() - ref_i) / ref_i)
    relb = ((dbb.double() - ref_b) / ref_b)
    groups = T * ((B + 31) // 32)
    print(f"B={B} T={T}: groups/CTA={groups / 148:.0f} MMAs/acc={groups / 148 * 12:.0f}  dW_ih rel err mean {rel.mean().item():+.3e} "
          f"max|.| {rel.abs().max().item():.3e};  db rel err mean {relb.mean().item():+.3e}")
