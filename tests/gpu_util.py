"""Helpers shared by the -m gpu parity tests: build the CUDA model from oracle/golden parameters, run a step."""
import torch

import b200asr
from oracle import asr_oracle as O

TOL = 1e-3      # north_star: "within 1e-3 relative fp32" (max|a-b| / max|b|)


def asr_cfg(ocfg: O.OracleConfig, dropout=0.0, smoothing=0.1) -> "b200asr.ASRConfig":
    return b200asr.ASRConfig(num_layers=ocfg.num_layers, num_heads=ocfg.num_heads, dim_model=ocfg.dim_model,
                             dim_key=ocfg.dim_key, dim_value=ocfg.dim_value, dim_inner=ocfg.dim_inner, vocab=ocfg.vocab,
                             feat_extractor=ocfg.feat_extractor, tgt_max_len=ocfg.tgt_max_len, src_max_len=ocfg.src_max_len,
                             freq=ocfg.freq, dropout=dropout, label_smoothing=smoothing, emb_trg_sharing=ocfg.emb_trg_sharing)


def cuda_model(ocfg, P, dropout=0.0, train=True):
    m = b200asr.build_model(asr_cfg(ocfg, dropout))
    missing, unexpected = m.load_state_dict({k: v.clone() for k, v in P.items()}, strict=False)
    assert not unexpected, unexpected
    assert all(k.endswith("positional_encoding.pe") or "running_" in k or "num_batches" in k for k in missing), missing
    m = m.cuda()
    m.train(train)
    return m


def cuda_step(model, src, lengths, tgt, smoothing):
    """forward + calculate_metrics + backward on the GPU; returns CPU copies."""
    model.zero_grad(set_to_none=True)
    pred, gold, hyp, _ = model(src.cuda(), lengths, tgt.cuda())
    loss, stats = b200asr.loss_and_stats(pred, gold, smoothing)
    loss.backward()
    torch.cuda.synchronize()
    grads = {n: (p.grad.detach().cpu() if p.grad is not None else torch.zeros_like(p).cpu()) for n, p in model.named_parameters()}
    return pred.detach().cpu(), gold.cpu(), hyp.cpu(), loss.detach().cpu(), stats.cpu(), grads


def cuda_step_with_decisions(model, src, lengths, tgt, smoothing):
    """cuda_step plus the path's own discrete decisions in the oracle's site naming (oracle.Decisions): which ReLU / Hardtanh
    units are active and which element every max-pool window takes, derived from the activations the kernels produced."""
    import importlib
    import torch.nn.functional as F
    ops = importlib.import_module(b200asr.__name__ + ".ops")
    ops.decision_capture = []
    try:
        out = cuda_step(model, src, lengths, tgt, smoothing)
        cap = ops.decision_capture
    finally:
        ops.decision_capture = None
    masks, n = {}, {}
    for kind, t in cap:
        i = n.get(kind, 0)
        n[kind] = i + 1
        if kind == "relu":
            masks[f"relu.{i}"] = (t > 0).cpu()
        elif kind == "hardtanh":
            masks[f"hardtanh.{i}"] = ((t > 0).to(torch.int8) + (t >= 20.0).to(torch.int8)).cpu()
        else:                                   # first maximum of the window, as ATen (and the pooling kernel) pick it
            masks[f"pool.{i}"] = F.max_pool2d(t.contiguous(), 2, 2, return_indices=True)[1].cpu()
    return out, masks
