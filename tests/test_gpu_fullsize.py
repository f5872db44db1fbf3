"""Full BASELINE cfg2 size (B=32, T=800, 161 bins, 4L/8H/d512, V=4364) on the GPU: the oracle cannot run this in seconds,
so the checks are size-independent properties of the path -- utterances are independent, hence
  * permuting the batch permutes `pred` (every kernel's batch indexing, masks and tile decomposition at full size), and
  * the un-normalised loss gradient of the batch is the sum of the gradients of its shards (what the data-parallel step
    relies on; covers every backward kernel incl. split-K / atomic accumulation at full size)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(dropout=0.0):
    import b200asr
    spec = b200asr.BASELINE_CONFIGS["cfg2"]
    cfg = spec["cfg"]
    import dataclasses
    cfg = dataclasses.replace(cfg, dropout=dropout)
    torch.manual_seed(1234)
    model = b200asr.build_model(cfg).cuda().train()
    B, T = spec["batch"], spec["t_src"]
    g = torch.Generator().manual_seed(7)
    src = torch.randn(B, 1, cfg.freq, T, generator=g).cuda()
    lens = torch.randint(T // 2, T + 1, (B,), generator=g).sort(descending=True).values.to(torch.int32)
    lens[0] = T
    for b in range(B):
        src[b, :, :, int(lens[b]):] = 0                               # zero-padded tails, as _collate_fn
    tgt = torch.randint(3, cfg.vocab, (B, cfg.tgt_max_len - 1), generator=g)
    for b in range(B):
        tgt[b, 20 + (5 * b) % 70:] = 0
    return b200asr, cfg, model, src, lens, tgt.cuda()


def _rel_l2(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def test_batch_permutation_permutes_predictions_at_cfg2_size():
    b200asr, cfg, model, src, lens, tgt = _setup()
    with torch.no_grad():
        pred, gold, hyp, _ = model(src, lens, tgt)
        perm = torch.randperm(src.shape[0], generator=torch.Generator().manual_seed(3))
        # the encoder masks come from the lengths, so a permuted batch is a different launch geometry for every kernel
        pred_p, gold_p, hyp_p, _ = model(src[perm.cuda()], lens[perm], tgt[perm.cuda()])
    assert torch.equal(gold_p, gold[perm.cuda()])
    assert _rel_l2(pred_p, pred[perm.cuda()]) < 2e-5
    real = gold.ne(0)
    agree = (hyp_p == hyp[perm.cuda()])[real[perm.cuda()]].float().mean().item()
    assert agree > 0.999                                              # argmax ids (near-ties at random init may flip)


def test_gradient_of_batch_is_sum_of_shard_gradients_at_cfg2_size():
    b200asr, cfg, model, src, lens, tgt = _setup()
    flat = b200asr.FlatParams(model)

    def grad_of(sl):
        flat.zero_grad()
        pred, gold, _, _ = model(src[sl], lens[sl], tgt[sl])
        loss_sum, stats = b200asr.loss_and_stats(pred, gold, cfg.label_smoothing, reduction="sum")
        loss_sum.backward()
        flat.ensure_grad_views()
        return flat.flat_grad[:flat.numel].clone(), float(loss_sum.detach()), int(stats[1])

    g_all, l_all, n_all = grad_of(slice(0, 32))
    g_a, l_a, n_a = grad_of(slice(0, 16))
    g_b, l_b, n_b = grad_of(slice(16, 32))
    assert n_all == n_a + n_b and abs(l_all - (l_a + l_b)) < 1e-5 * abs(l_all)
    # summation order differs between the runs (token ranges of the split-K / atomic accumulations), nothing else:
    # measured 2.5e-4 in the global norm; the bar is the path's 1e-3
    assert _rel_l2(g_a + g_b, g_all) < 1e-3
    # per-parameter check on the big tensors (a single mis-tiled kernel would hide in the global norm; the small ones
    # include the key biases, whose gradient is mathematically zero, i.e. pure rounding noise)
    for p, o in zip(flat.params, flat.offsets):
        n = p.numel()
        if n >= 65536:
            assert _rel_l2((g_a + g_b)[o:o + n], g_all[o:o + n]) < 1e-3
