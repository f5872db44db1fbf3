#!/usr/bin/env python
"""Generate golden fixtures by importing the UNMODIFIED reference from /root/reference.

Run in the build container only (the GPU box has no reference checkout):

    python tests/golden/make_golden.py            # regenerates every case

Each case runs in its own interpreter because the reference parses its flags at
import time (utils/constant.py:99).  Shims (none edit the reference):
  * sys.modules['Levenshtein'] stub -- python-Levenshtein is not installed and is only
    called by calculate_cer/wer, which are off the hot path (utils/metrics.py:3,56,76);
  * sys.argv set before the first `import utils.constant`.
A fixture holds: the inputs, the full state_dict (reference names), pred / gold /
hyp_seq of Transformer.forward (models/asr/transformer.py:59-85), the loss and
num_correct of calculate_metrics (utils/metrics.py:78-95) and every parameter gradient
after loss.backward() -- all fp32 from torch CPU kernels.
"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"

CASES = {
    # name: (flags, B, T_src, freq, V, smoothing, ragged)
    "none_l1": (["--num-layers", "1", "--num-heads", "2", "--dim-model", "128", "--dim-emb", "128",
                 "--dim-key", "64", "--dim-value", "64", "--dim-inner", "256", "--feat_extractor", "",
                 "--tgt-max-len", "10"], 2, 50, 161, 32, 0.1, True),
    "none_ce": (["--num-layers", "2", "--num-heads", "4", "--dim-model", "64", "--dim-emb", "64",
                 "--dim-key", "16", "--dim-value", "32", "--dim-inner", "96", "--feat_extractor", "",
                 "--tgt-max-len", "12"], 3, 37, 161, 50, 0.0, True),
    "vgg_l2": (["--num-layers", "2", "--num-heads", "2", "--dim-model", "64", "--dim-emb", "64",
                "--dim-key", "32", "--dim-value", "32", "--dim-inner", "128", "--feat_extractor", "vgg_cnn",
                "--sample-rate", "4000", "--tgt-max-len", "8"], 2, 24, 41, 40, 0.1, True),
    "emb_l1": (["--num-layers", "1", "--num-heads", "2", "--dim-model", "64", "--dim-emb", "64",
                "--dim-key", "32", "--dim-value", "32", "--dim-inner", "128", "--feat_extractor", "emb_cnn",
                "--tgt-max-len", "8"], 3, 40, 161, 40, 0.1, False),
}


def run_case(name):
    import types
    flags, B, T, freq, V, smoothing, ragged = CASES[name]
    sys.path.insert(0, REF)
    sys.modules["Levenshtein"] = types.ModuleType("Levenshtein")
    sys.argv = ["make_golden"] + flags + ["--label-smoothing", str(smoothing), "--dropout", "0.0"]
    import torch
    from utils import constant
    from utils.functions import init_transformer_model
    from utils.metrics import calculate_metrics

    labels = ["_", "<s>", "</s>"] + ["t%d" % i for i in range(V - 3)]
    label2id = {l: i for i, l in enumerate(labels)}
    id2label = {i: l for i, l in enumerate(labels)}
    torch.manual_seed(123456)
    model = init_transformer_model(constant.args, label2id, id2label)
    model.train()
    # nudge norm/bias parameters away from (1, 0) so their gradients and use are exercised
    g = torch.Generator().manual_seed(7)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn(p.shape, generator=g))

    g = torch.Generator().manual_seed(0)
    src = torch.randn(B, 1, freq, T, generator=g)
    if ragged:
        lens = (T * (0.5 + 0.5 * torch.rand(B, generator=g))).long().clamp(1, T)
        lens[0] = T
        lens, _ = torch.sort(lens, descending=True)
    else:
        lens = torch.full((B,), T, dtype=torch.long)
    for i in range(B):
        src[i, :, :, int(lens[i]):] = 0
    L = constant.args.tgt_max_len - 1
    tgt = torch.randint(3, V, (B, L), generator=g)
    tl = (L * (0.4 + 0.6 * torch.rand(B, generator=g))).long().clamp(1, L)
    tl[0] = L
    for i in range(B):
        tgt[i, int(tl[i]):] = 0
    lens32 = lens.to(torch.int32)

    pred, gold, hyp, _ = model(src, lens32, tgt, verbose=False)
    loss, ncorrect = calculate_metrics(pred, gold, smoothing=smoothing, loss_type="ce")
    loss.backward()

    out = {"in.src": src.numpy(), "in.lengths": lens32.numpy(), "in.tgt": tgt.numpy(),
           "out.pred": pred.detach().numpy(), "out.gold": gold.numpy(), "out.hyp": hyp.numpy(),
           "out.loss": np.float32(loss.item()), "out.num_correct": np.int64(ncorrect),
           "meta.smoothing": np.float32(smoothing)}
    for n, p in model.state_dict().items():
        if n.endswith("positional_encoding.pe") or "num_batches_tracked" in n or "running_" in n:
            continue
        out["param." + n] = p.detach().numpy()
    for n, p in model.named_parameters():
        out["grad." + n] = p.grad.numpy() if p.grad is not None else np.zeros(p.shape, np.float32)
    a = constant.args
    out["meta.cfg"] = np.array([a.num_layers, a.num_heads, a.dim_model, a.dim_key, a.dim_value, a.dim_inner,
                                V, a.tgt_max_len, freq], dtype=np.int64)
    out["meta.feat"] = np.array(a.feat_extractor)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "pred", tuple(pred.shape), "loss %.6f" % loss.item(), "->", path,
          "%.1f KB" % (os.path.getsize(path) / 1024))


def run_greedy():
    """Fixture `greedy`: token ids of the reference's Decoder.greedy_search (models/asr/transformer.py:316-394) -- 300 fixed
    steps of full-prefix re-decode, strings cut at the first EOS (:385-393).  Shims (none edit the
    reference): get_subsequent_mask(...).bool() (the uint8 mask no longer adds to a bool one on torch >= 1.2) and
    --tgt-max-len 301 (the positional table must cover 300 positions).  Every id >= 3 maps to its own character, so
    the returned strings map back to ids.  The EOS row of output_linear is scaled so that utterances end at different,
    non-trivial steps (a random-init model would otherwise hardly ever emit EOS)."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import ref_shim
    V, B, Te = 40, 4, 23
    ns = ref_shim.load(["--num-layers", "2", "--num-heads", "2", "--dim-model", "64", "--dim-emb", "64", "--dim-key", "32",
                        "--dim-value", "32", "--dim-inner", "128", "--feat_extractor", "", "--tgt-max-len", "301",
                        "--dropout", "0.0"], q12_shim=True)
    import torch
    l2i, i2l = ref_shim.labels(V)
    torch.manual_seed(123456)
    model = ns.functions.init_transformer_model(ns.constant.args, l2i, i2l)
    model.eval()
    g = torch.Generator().manual_seed(11)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn(p.shape, generator=g))
        model.decoder.output_linear.weight.mul_(4.0)            # spread the logits (fewer near-ties)
        model.decoder.output_linear.weight[2].mul_(2.2)         # EOS wins now and then
    enc = torch.randn(B, Te, 64, generator=g)
    with torch.no_grad():
        strs = model.decoder.greedy_search(enc)
    ids = np.full((B, 300), -1, dtype=np.int64)                 # -1 = not produced (after the EOS cut)
    for b, s in enumerate(strs):
        ids[b, :len(s)] = [l2i[c] for c in s]
    out = {"in.enc": enc.numpy(), "out.ids": ids, "out.lengths": np.array([len(s) for s in strs], dtype=np.int64)}
    for n, p in model.state_dict().items():
        if n.startswith("decoder.") and not n.endswith("positional_encoding.pe"):
            out["param." + n] = p.detach().numpy()
    out["meta.cfg"] = np.array([2, 2, 64, 32, 32, 128, V, 301, 161], dtype=np.int64)
    path = os.path.join(HERE, "greedy.npz")
    np.savez_compressed(path, **out)
    print("greedy: utterance lengths (cut at EOS)", [len(s) for s in strs], "->", path, "%.1f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    if len(sys.argv) == 3 and sys.argv[1] == "--case":
        run_greedy() if sys.argv[2] == "greedy" else run_case(sys.argv[2])
    else:
        for c in list(CASES) + ["greedy"]:
            subprocess.check_call([sys.executable, "-W", "ignore", os.path.abspath(__file__), "--case", c])
