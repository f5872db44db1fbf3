#!/usr/bin/env python
"""N training steps of the bench workload (cfg2 by default) with nothing else around them -- the command ncu wraps.
usage: one_step.py [steps] [cfg] [batch]"""
import sys

import torch

sys.path.insert(0, ".")
import b200asr  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
name = sys.argv[2] if len(sys.argv) > 2 else "cfg2"
spec = b200asr.BASELINE_CONFIGS[name]
cfg, B, T = spec["cfg"], (int(sys.argv[3]) if len(sys.argv) > 3 else spec["batch"]), spec["t_src"]
dev = torch.device("cuda")
torch.manual_seed(123456)
model = b200asr.build_model(cfg).to(dev).train()
dp = b200asr.DataParallelStep(model, model_size=cfg.dim_input, smoothing=cfg.label_smoothing)
g = torch.Generator().manual_seed(0)
src = torch.randn(B, 1, cfg.freq, T, generator=g).to(dev)
tgt = torch.randint(3, cfg.vocab, (B, cfg.tgt_max_len - 1), generator=g).to(dev)
lens = torch.full((B,), T, dtype=torch.int32)
lib = b200asr._lib.load(check_device=True)
for i in range(steps):
    n0 = lib.b200asr_launch_count()
    dp.step(src, lens, tgt)
    torch.cuda.synchronize()
    print(f"step {i}: {lib.b200asr_launch_count() - n0} launches, loss {float(dp.global_loss()):.4f}", flush=True)
