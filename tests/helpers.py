"""Shared helpers for the tests (fixtures loading, tolerances)."""
import os

import numpy as np
import torch

from oracle import asr_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CASES = ["none_l1", "none_ce", "vgg_l2", "emb_l1"]


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    c = z["meta.cfg"]
    cfg = O.OracleConfig(num_layers=int(c[0]), num_heads=int(c[1]), dim_model=int(c[2]), dim_key=int(c[3]),
                         dim_value=int(c[4]), dim_inner=int(c[5]), vocab=int(c[6]), tgt_max_len=int(c[7]),
                         freq=int(c[8]), feat_extractor=str(z["meta.feat"]))
    P = {k[len("param."):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param.")}
    G = {k[len("grad."):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("grad.")}
    io = {k: torch.from_numpy(np.asarray(z[k])) for k in z.files if k.startswith("in.") or k.startswith("out.")}
    return cfg, P, G, io, float(z["meta.smoothing"])


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max|a-b| / max|b| -- the 'relative fp32' measure used for every parity statement."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    denom = b.abs().max().item()
    if denom == 0.0:
        return (a - b).abs().max().item()
    return (a - b).abs().max().item() / denom


def grads_rel_err(grads: dict, ref: dict) -> dict:
    """Per-tensor max|g-g_ref| / max(max|g_ref|, 1e-3 * largest gradient magnitude in the model).
    The floor keeps mathematically-zero gradients (e.g. key_linear.bias: softmax is shift invariant)
    from turning rounding noise into a relative error of O(1)."""
    gmax = max(float(v.abs().max()) for v in ref.values())
    out = {}
    for k, r in ref.items():
        denom = max(float(r.abs().max()), 1e-3 * gmax)
        out[k] = float((grads[k].detach().double().cpu() - r.double().cpu()).abs().max()) / denom
    return out
