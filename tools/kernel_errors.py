#!/usr/bin/env python
"""Per-kernel arithmetic error against float64 at cfg2-like shapes, every precision mode: linear fwd / dgrad / wgrad,
attention fwd / bwd.  max|err| / max|ref| (and the relative L2 error in brackets).  GPU only."""
import importlib
import sys

import torch

sys.path.insert(0, ".")
import b200asr  # noqa: E402

ops = importlib.import_module(b200asr.__name__ + ".ops")
L = importlib.import_module(b200asr.__name__ + "._lib")
torch.manual_seed(0)


def err(a, b):
    a, b = a.double(), b.double()
    return "%.1e [%.1e]" % (float((a - b).abs().max() / b.abs().max()), float((a - b).norm() / b.norm()))


only_attn = "--attn" in sys.argv
print("---- linear: y = x w^T + b, dx = dy w, dw = dy^T x, db")
for (M, N, K) in [] if only_attn else [(400, 512, 512), (800, 1536, 512), (6400, 2048, 512), (6400, 512, 2048), (400, 4364, 512), (800, 512, 5120)]:
    x = torch.randn(M, K, device="cuda")
    w = torch.randn(N, K, device="cuda") * K ** -0.5
    b = torch.randn(N, device="cuda")
    dy = torch.randn(M, N, device="cuda")
    y64 = x.double() @ w.double().t() + b.double()
    dx64 = dy.double() @ w.double()
    dw64 = dy.double().t() @ x.double()
    db64 = dy.double().sum(0)
    for prec in (0, 1, 3, 2, 6):
        ws = ops.split_weight(w, prec)
        y = ops.linear_fwd(x, w, b, False, prec, ws)
        dx = ops.linear_bwd_data(dy, w, None, prec, ws)
        dw, db = ops.linear_bwd_weight(dy, x, True, prec)
        print(f"M={M:5d} N={N:5d} K={K:5d} p{prec}: fwd {err(y, y64)}  dgrad {err(dx, dx64)}  wgrad {err(dw, dw64)}  dbias {err(db, db64)}", flush=True)

print("---- attention (B=4, H=8): out, dq, dk, dv")
for (Tq, Tk, causal) in [(200, 200, False), (100, 100, True), (100, 200, False), (250, 250, False), (400, 400, False), (37, 150, False)]:
    B, H, d = 4, 8, 64
    q = torch.randn(B, Tq, H, d, device="cuda").permute(0, 2, 1, 3)
    k = torch.randn(B, Tk, H, d, device="cuda").permute(0, 2, 1, 3)
    v = torch.randn(B, Tk, H, d, device="cuda").permute(0, 2, 1, 3)
    do = torch.randn(B, Tq, H, d, device="cuda").permute(0, 2, 1, 3)
    for scale_name, qs in (("unit", 1.0), ("flat", 0.05)):          # flat: near-uniform attention as at random init
        q64 = (q * qs).double().requires_grad_(True); k64 = k.double().requires_grad_(True); v64 = v.double().requires_grad_(True)
        s = q64 @ k64.transpose(2, 3) / 8.0
        if causal:
            s = s.masked_fill(torch.triu(torch.ones(Tq, Tk, device="cuda", dtype=torch.bool), 1), float("-inf"))
        o64 = torch.softmax(s, -1) @ v64
        o64.backward(do.double())
        for attn in ("fp32", "tf32x3", "bf16x3", "tf32"):
            ops.config.set(attn=attn, attn_bwd="tf32" if attn == "tf32" else "fp32")
            qq = (q * qs).detach().clone().requires_grad_(True); kk = k.detach().clone().requires_grad_(True); vv = v.detach().clone().requires_grad_(True)
            o = ops.SdpaFn.apply(qq, kk, vv, None, None, causal, 1.0 / 8.0, 0.0)
            o.backward(do)
            print(f"Tq={Tq} Tk={Tk} causal={int(causal)} {scale_name:4s} {attn:6s}: out {err(o, o64)}  dq {err(qq.grad, q64.grad)}  dk {err(kk.grad, k64.grad)}  dv {err(vv.grad, v64.grad)}", flush=True)
