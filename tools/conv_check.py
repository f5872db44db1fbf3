#!/usr/bin/env python
"""GPU check of the 3x3 convolution forward / data gradient in every tensor-core mode against float64 (cuDNN-free: torch
conv2d in double on the GPU)."""
import importlib
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
import b200asr  # noqa: E402

ops = importlib.import_module(b200asr.__name__ + ".ops")
L = importlib.import_module(b200asr.__name__ + "._lib")
lib = L.load(check_device=True)
st = torch.cuda.current_stream().cuda_stream
torch.manual_seed(0)
for (B, T, Fq, Ci, Co) in [(1, 16, 8, 32, 64), (2, 37, 21, 64, 64), (2, 40, 40, 64, 128), (1, 50, 20, 128, 128)]:
    x = torch.randn(B, T, Fq, Ci, device="cuda")                   # channels-last [B,T,F,C]
    w = torch.randn(Co, Ci, 3, 3, device="cuda") * 0.1             # (Co, Ci, kF, kT) as nn.Conv2d over (F, T)
    bias = torch.randn(Co, device="cuda")
    dy = torch.randn(B, T, Fq, Co, device="cuda")
    xn = x.permute(0, 3, 2, 1).double()                            # (B, C, F, T)
    y64 = F.conv2d(xn, w.double(), bias.double(), padding=1).permute(0, 3, 2, 1)
    dx64 = torch.autograd.grad(F.conv2d(xn.requires_grad_(True), w.double(), None, padding=1), xn, dy.permute(0, 3, 2, 1).double())[0].permute(0, 3, 2, 1)
    xg = xn.detach().clone().requires_grad_(False)
    w64 = w.double().requires_grad_(True)
    b64 = bias.double().requires_grad_(True)
    F.conv2d(xg, w64, b64, padding=1).backward(dy.permute(0, 3, 2, 1).double())
    dw64, db64 = w64.grad, b64.grad
    ws = torch.empty(lib.b200asr_conv3x3_ws_bytes(max(Ci, Co), max(Ci, Co)) // 4, device="cuda")
    line = f"B={B} T={T} F={Fq} Ci={Ci} Co={Co} |"
    for prec in (0, 1, 3, 2, 6):
        y = torch.empty(B, T, Fq, Co, device="cuda")
        dx = torch.empty(B, T, Fq, Ci, device="cuda")
        rc1 = lib.b200asr_conv3x3_fwd(x.data_ptr(), w.data_ptr(), bias.data_ptr(), y.data_ptr(), ws.data_ptr(), B, T, Fq, Ci, Co, 0, prec, st)
        dx16 = torch.empty(2, B, T, Fq, Ci, device="cuda", dtype=torch.bfloat16) if prec in (2, 6) else None
        rc2 = lib.b200asr_conv3x3_bwd_data(dy.data_ptr(), w.data_ptr(), None, dx.data_ptr(), dx16.data_ptr() if dx16 is not None else None, ws.data_ptr(), B, T, Fq, Ci, Co, prec, st)
        torch.cuda.synchronize()
        if rc1 or rc2:
            line += f" p{prec}: rc {rc1},{rc2} {L.last_error()[:60]} |"
            continue
        ey = float((y.double() - y64).abs().max() / y64.abs().max())
        ex = float((dx.double() - dx64).abs().max() / dx64.abs().max())
        line += f" p{prec}: fwd {ey:.1e} dgrad {ex:.1e}"
        if dx16 is not None:
            line += f" dx16 {float(((dx16[0].double() + dx16[1].double()) - dx64).abs().max() / dx64.abs().max()):.1e}"
        if Ci % 64 == 0:
            dw = torch.empty_like(w)
            db = torch.empty_like(bias)
            rc3 = lib.b200asr_conv3x3_bwd_weight(dy.data_ptr(), None, x.data_ptr(), dw.data_ptr(), db.data_ptr(), ws.data_ptr(), B, T, Fq, Ci, Co, prec, st)
            torch.cuda.synchronize()
            if rc3:
                line += f" wgrad rc {rc3} {L.last_error()[:50]}"
            else:
                line += f" wgrad {float((dw.double() - dw64).abs().max() / dw64.abs().max()):.1e} db {float((db.double() - db64).abs().max() / db64.abs().max()):.1e}"
            if prec in (2, 6):      # the same with dy arriving as bf16 pairs (B tiles by TMA)
                hi = dy.bfloat16()
                dy16 = torch.stack([hi, (dy - hi.float()).bfloat16()]).contiguous()
                dw2 = torch.empty_like(w); db2 = torch.empty_like(bias)
                rc4 = lib.b200asr_conv3x3_bwd_weight(dy.data_ptr(), dy16.data_ptr(), x.data_ptr(), dw2.data_ptr(), db2.data_ptr(), ws.data_ptr(), B, T, Fq, Ci, Co, prec, st)
                torch.cuda.synchronize()
                line += f" pairs: rc {rc4} wgrad {float((dw2.double() - dw64).abs().max() / dw64.abs().max()):.1e} db {float((db2.double() - db64).abs().max() / db64.abs().max()):.1e}"
        line += " |"
    print(line, flush=True)
