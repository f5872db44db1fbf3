#!/usr/bin/env python
"""One training step (fwd + CE + bwd + Adam) at every BASELINE.json configuration's per-GPU shape: checks that each
shape is supported end to end (finite loss, no shape/alignment rejection) and prints ms/step.  Usage:
    python tools/run_configs.py [cfg1 cfg2 ...] [--batch N]"""
import sys
import time

import torch

sys.path.insert(0, ".")
import b200asr  # noqa: E402

names = [a for a in sys.argv[1:] if a.startswith("cfg")] or list(b200asr.BASELINE_CONFIGS)
batch_override = int(sys.argv[sys.argv.index("--batch") + 1]) if "--batch" in sys.argv else 0
dev = torch.device("cuda")
for name in names:
    spec = b200asr.BASELINE_CONFIGS[name]
    cfg, B, T = spec["cfg"], batch_override or spec["batch"], spec["t_src"]
    torch.manual_seed(0)
    model = b200asr.build_model(cfg).to(dev).train()
    dp = b200asr.DataParallelStep(model, model_size=cfg.dim_input, smoothing=cfg.label_smoothing)
    src = torch.randn(B, 1, cfg.freq, T, device=dev)
    lens = torch.full((B,), T, dtype=torch.int32)
    tgt = torch.randint(3, cfg.vocab, (B, cfg.tgt_max_len - 1), device=dev)
    try:
        dp.step(src, lens, tgt)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(2):
            dp.step(src, lens, tgt)
        loss = float(dp.global_loss())
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 500
        print(f"{name}: B={B} T={T} T_enc={cfg.t_enc(T)} L={cfg.num_layers} d={cfg.dim_model} feat={cfg.feat_extractor or 'none'} "
              f"params={sum(p.numel() for p in model.parameters())/1e6:.1f}M  {ms:.1f} ms/step  {B/ms*1e3:.0f} utt/s  loss {loss:.4f} "
              f"mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB")
        assert loss == loss and abs(loss) < 1e4
    except Exception as e:  # noqa: BLE001
        print(f"{name}: FAILED {type(e).__name__}: {str(e)[:300]}")
    del model, dp
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()
