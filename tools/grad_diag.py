#!/usr/bin/env python
"""Per-tensor gradient errors of the CUDA path against the fp64 oracle at a true BASELINE shape (reduced batch), for the
exact-fp32 build and the default; localises which tensors carry the error.  usage: grad_diag.py [cfg] [batch] [mode ...]"""
import importlib
import sys

import torch

sys.path.insert(0, ".")
import b200asr  # noqa: E402
from oracle import asr_oracle as O  # noqa: E402
from tests.gpu_util import cuda_model, cuda_step, cuda_step_with_decisions  # noqa: E402
from tests.helpers import grads_rel_err, rel_err  # noqa: E402

ops = importlib.import_module(b200asr.__name__ + ".ops")
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
modes = sys.argv[3:] or ["fp32", "default"]
spec = b200asr.BASELINE_CONFIGS[name]
c = spec["cfg"]
ocfg = O.OracleConfig(num_layers=c.num_layers, num_heads=c.num_heads, dim_model=c.dim_model, dim_key=c.dim_key, dim_value=c.dim_value,
                      dim_inner=c.dim_inner, vocab=c.vocab, feat_extractor=c.feat_extractor, tgt_max_len=c.tgt_max_len, freq=c.freq)
P = O.init_params(ocfg, seed=123456)
g = torch.Generator().manual_seed(5)
for k, v in P.items():
    if v.dim() == 1:
        v.add_(0.1 * torch.randn(v.shape, generator=g))
src, lens, tgt = O.synthetic_batch(ocfg, B, spec["t_src"], seed=0, ragged=True)
torch.set_num_threads(32)
rec = O.Decisions()
P64 = {k: v.double() for k, v in P.items()}
r64 = O.forward_backward(P64, ocfg, src.double(), lens, tgt, 0.1, dec=rec)
r32 = O.forward_backward(P, ocfg, src, lens, tgt, 0.1)
g64 = {k: v.float() for k, v in r64[5].items()}
e32 = grads_rel_err(r32[5], g64)
print("fp32 oracle vs fp64: pred %.2e grads max %.2e median %.2e" % (rel_err(r32[0], r64[0]), max(e32.values()), sorted(e32.values())[len(e32) // 2]))
defaults = (ops.config.linear, ops.config.conv, ops.config.attn, ops.config.conv_wgrad, ops.config.attn_bwd)
for mode in modes:
    ops.config.linear, ops.config.conv, ops.config.attn, ops.config.conv_wgrad, ops.config.attn_bwd = defaults
    if mode == "default":
        pass
    elif "," in mode:
        names = ["linear", "conv", "attn", "conv_wgrad", "attn_bwd"]
        ops.config.set(**dict(zip(names, mode.split(","))))
    else:
        ops.config.set(linear=mode, conv=mode, attn=mode if mode in ("fp32", "tf32", "tf32x3") else "tf32x3", conv_wgrad=mode, attn_bwd="fp32")
    model = cuda_model(ocfg, P)
    (pred, gold, hyp, loss, stats, grads), masks = cuda_step_with_decisions(model, src, lens, tgt, 0.1)
    # decisions that differ from the fp64 oracle's own, and how close to their threshold those units are in fp64
    for site in sorted(masks):
        diff = masks[site] != rec.masks[site]
        nd = int(diff.sum())
        if nd:
            pre = rec.pre[site]
            print(f"   {site:12s} {nd:7d} of {diff.numel():10d} decisions differ; |pre-activation| / max at those: max {float(pre[diff].abs().max() / pre.abs().max()):.1e}")
    r64f = O.forward_backward(P64, ocfg, src.double(), lens, tgt, 0.1, dec=O.Decisions(frozen=masks))
    g64f = {k: v.float() for k, v in r64f[5].items()}
    errs_unfrozen = grads_rel_err(grads, g64)
    print(f"   unfrozen: grads max {max(errs_unfrozen.values()):.2e} median {sorted(errs_unfrozen.values())[len(errs_unfrozen) // 2]:.2e}; below: against the fp64 oracle evaluated AT THE PATH'S OWN DECISIONS")
    g64_saved, g64 = g64, g64f
    errs = grads_rel_err(grads, g64)
    order = sorted(errs, key=errs.get, reverse=True)
    print(f"== {mode}: pred {rel_err(pred, r64[0]):.2e} loss {abs(loss.item() - r64[3].item()) / abs(r64[3].item()):.2e} "
          f"grads max {errs[order[0]]:.2e} median {errs[order[len(order) // 2]]:.2e}  #>1e-3: {sum(e > 1e-3 for e in errs.values())}/{len(errs)}")
    for k in order[:12]:
        r, x = g64[k], grads[k].float()
        l2 = float((x - r).norm() / r.norm().clamp_min(1e-30))
        rows = ""
        if r.dim() >= 2:      # are the errors confined to a few rows (= flipped units) or spread over the tensor?
            d = (x - r).reshape(r.shape[0], -1).abs().max(1).values / r.abs().max()
            rows = f" rows>1e-3: {int((d > 1e-3).sum())}/{d.numel()} top {[round(float(v), 4) for v in d.topk(min(3, d.numel())).values]} row-median {float(d.median()):.1e}"
        print(f"   {k:55s} max-rel {errs[k]:.2e}  rel-L2 {l2:.2e}  max|ref| {float(r.abs().max()):.2e}  fp32-oracle {e32[k]:.2e}{rows}")
    print("   best:", ", ".join(f"{k.split('.')[-3:]} {errs[k]:.1e}" for k in order[-4:]))
    ratio = {k: errs[k] / max(e32[k], 1e-7) for k in errs}
    kr = max(ratio, key=ratio.get)
    print(f"   error / fp32-oracle error: max {ratio[kr]:.0f} ({kr}) median {sorted(ratio.values())[len(ratio) // 2]:.0f}")
    g64 = g64_saved
