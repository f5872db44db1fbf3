#!/usr/bin/env python
"""End-to-end parity of the CUDA path against the CPU oracle at the cfg2 ARCHITECTURE (4L/8H/d512/vgg, V=4364) on a
reduced batch (B=2, T_src=200, T_tgt=30) for several precision mixes.  Prints max-relative errors (pred, loss, grads)."""
import importlib
import sys
import time

import torch

sys.path.insert(0, ".")
import b200asr  # noqa: E402
from oracle import asr_oracle as O  # noqa: E402
from tests.gpu_util import cuda_model, cuda_step  # noqa: E402
from tests.helpers import grads_rel_err, rel_err  # noqa: E402

ops = importlib.import_module(b200asr.__name__ + ".ops")
B, T, Tt = (int(a) for a in (sys.argv[1:4] + ["2", "200", "30"])[:3])
cfg = O.OracleConfig(num_layers=4, num_heads=8, dim_model=512, dim_key=64, dim_value=64, dim_inner=2048, vocab=4364,
                     feat_extractor="vgg_cnn", tgt_max_len=Tt, freq=161)
P = O.init_params(cfg, seed=123456)
src, lens, tgt = O.synthetic_batch(cfg, B, T, seed=0, ragged=True)
t0 = time.time()
Pd = {k: v.double() for k, v in P.items()}
pred64, gold, hyp64, loss64, n_word, grads64 = O.forward_backward(Pd, cfg, src.double(), lens, tgt, 0.1)
pred32, _, hyp32, loss32, _, grads32 = O.forward_backward(P, cfg, src, lens, tgt, 0.1)
print("oracle fp64+fp32 %.1fs; fp32 oracle vs fp64: pred %.2e loss %.2e grads %.2e" % (
    time.time() - t0, rel_err(pred32, pred64), abs(loss32.item() - loss64.item()) / abs(loss64.item()),
    max(grads_rel_err(grads32, {k: v.float() for k, v in grads64.items()}).values())))
real = gold.ne(O.PAD)
model = cuda_model(cfg, P)
mixes = [("fp32", "fp32", "fp32", "fp32", "fp32"), ("tf32x3", "tf32x3", "tf32x3", "tf32x3", "fp32"),   # 2nd = package default
         ("fp32", "fp32", "fp32", "tf32x3", "fp32"), ("tf32x3", "tf32x3", "tf32x3", "fp32", "fp32"),
         ("tf32x3", "tf32x3", "tf32x3", "tf32", "tf32"), ("tf32x3", "tf32x3", "tf32x3", "tf32", "fp32"),
         ("fp32", "fp32", "fp32", "tf32", "tf32"), ("tf32x3", "tf32", "tf32", "tf32", "tf32"), ("tf32", "tf32", "tf32", "tf32", "tf32")]
for lin, conv, convw, attn, attnb in mixes:
    ops.config.set(linear=lin, conv=conv, conv_wgrad=convw, attn=attn, attn_bwd=attnb)
    try:
        pred, g2, hyp, loss, stats, grads = cuda_step(model, src, lens, tgt, 0.1)
    except RuntimeError as e:
        print((lin, conv, convw, attn, attnb), "ERROR", str(e)[:150])
        continue
    errs = grads_rel_err(grads, {k: v.float() for k, v in grads64.items()})
    worst = max(errs, key=errs.get)
    flips = int((hyp[real] != hyp64[real]).sum())
    print("linear=%-6s conv=%-6s conv_wgrad=%-6s attn=%-4s/%-4s | pred %.2e loss %.2e grads max %.2e (%s) median %.2e argmax flips %d/%d" % (
        lin, conv, convw, attn, attnb, rel_err(pred, pred64), abs(loss.item() - loss64.item()) / abs(loss64.item()), errs[worst], worst,
        sorted(errs.values())[len(errs) // 2], flips, int(real.sum())))
