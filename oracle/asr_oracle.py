"""CPU oracle for the speech-Transformer training hot path (TEST INFRASTRUCTURE ONLY).

This file is a functional, CPU-only restatement of the algorithm that
gentaiscool/end2end-asr-pytorch runs for one teacher-forced forward pass, the
label-smoothed cross-entropy and (through torch autograd on these very ops) the
backward pass.  It exists so the CUDA path can be checked on a machine that does
not have the reference checkout (the GPU box).  It is *not* part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg
may import it.  The product package never does, and fails loudly without its
CUDA library.

Pinning: the reference ships no tests or golden vectors, so the
oracle is pinned against outputs of the reference itself, imported live in the
build container by ``tests/golden/make_golden.py``; the resulting fixtures are
committed under ``tests/golden/`` and ``tests/test_oracle_golden.py`` replays
them through this file.  Third-party arithmetic: everything below bottoms out
in PyTorch CPU kernels (reference README pins "Pytorch 1.4"; here torch 2.11).

Every function cites the reference lines (relative to /root/reference) that it
follows.  Parameters are passed as a flat ``dict`` keyed by the reference's own
``state_dict`` names so a reference checkpoint can be fed in unchanged.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

PAD, SOS, EOS = 0, 1, 2  # utils/constant.py:102-104


@dataclass
class OracleConfig:
    """Shape of one model instance (values of utils/constant.py:52-62 flags)."""
    num_layers: int = 4
    num_heads: int = 8
    dim_model: int = 512
    dim_key: int = 64
    dim_value: int = 64
    dim_inner: int = 2048
    vocab: int = 4364
    feat_extractor: str = "vgg_cnn"   # 'vgg_cnn' | 'emb_cnn' | ''
    tgt_max_len: int = 100
    src_max_len: int = 4000
    freq: int = 161
    emb_trg_sharing: bool = False
    extra: dict = field(default_factory=dict)

    @property
    def dim_input(self) -> int:
        # utils/functions.py:120-130
        if self.feat_extractor == "vgg_cnn":
            return (self.freq // 2 // 2) * 128
        if self.feat_extractor == "emb_cnn":
            h = (self.freq - 41) // 2 + 1
            h = (h - 21) // 2 + 1
            return h * 32
        return self.freq


def sinusoid_table(length: int, dim: int, dtype=torch.float32) -> torch.Tensor:
    """models/common_layers.py:83-87 -- sin on even columns, cos on odd."""
    pos = torch.arange(0, length, dtype=torch.float32).unsqueeze(1)
    freq = torch.exp(torch.arange(0, dim, 2, dtype=torch.float32) * -(math.log(10000.0) / dim))
    table = torch.zeros(length, dim, dtype=torch.float32)
    table[:, 0::2] = torch.sin(pos * freq)
    table[:, 1::2] = torch.cos(pos * freq)
    return table.to(dtype)


# --------------------------------------------------------------------------- discontinuities (ReLU / max-pool / Hardtanh)
class Decisions:
    """The model's discrete decisions -- which ReLU / Hardtanh units are active, which element each max-pool window takes.
    They are the only discontinuities of the path: between two arithmetics that agree to ~1e-5 a unit sitting within
    rounding of its threshold may decide differently, and its whole (O(1)) gradient contribution moves.  `record` mode
    stores the decisions (and the pre-activations) of a run; `frozen` mode replays given decisions, which makes the
    function smooth in its inputs -- the parity tests compare the CUDA gradients with the fp64 oracle evaluated AT THE
    CUDA PATH'S OWN DECISIONS, and separately count and bound the decisions that differ."""

    def __init__(self, frozen: Optional[Dict[str, torch.Tensor]] = None):
        self.frozen = frozen
        self.masks: Dict[str, torch.Tensor] = {}     # site -> bool mask (ReLU/Hardtanh: active) or int64 window indices (pool)
        self.pre: Dict[str, torch.Tensor] = {}       # site -> float32 pre-activation (ReLU) / top-2 window gap (pool)
        self._n = {}

    def _site(self, kind):
        i = self._n.get(kind, 0)
        self._n[kind] = i + 1
        return f"{kind}.{i}"

    def relu(self, x):
        site = self._site("relu")
        if self.frozen is not None:
            return x * self.frozen[site].to(x.dtype)
        self.masks[site] = x.detach() > 0
        self.pre[site] = x.detach().float()
        return F.relu(x)

    def hardtanh(self, x, lo=0.0, hi=20.0):
        site = self._site("hardtanh")
        if self.frozen is not None:
            m = self.frozen[site]                       # int8: 0 = clamped low, 1 = linear, 2 = clamped high
            return x * (m == 1).to(x.dtype) + lo * (m == 0).to(x.dtype) + hi * (m == 2).to(x.dtype)
        self.masks[site] = ((x.detach() > lo).to(torch.int8) + (x.detach() >= hi).to(torch.int8))
        self.pre[site] = torch.minimum((x.detach() - lo).abs(), (x.detach() - hi).abs()).float()
        return torch.clamp(x, lo, hi)

    def max_pool(self, x):
        site = self._site("pool")
        B, C, H, W = x.shape
        if self.frozen is not None:
            idx = self.frozen[site]
            return x.flatten(2).gather(2, idx.flatten(2)).view(idx.shape)
        y, idx = F.max_pool2d(x, 2, 2, return_indices=True)
        self.masks[site] = idx
        w = x.detach()[:, :, :H // 2 * 2, :W // 2 * 2].reshape(B, C, H // 2, 2, W // 2, 2).permute(0, 1, 2, 4, 3, 5).reshape(B, C, H // 2, W // 2, 4)
        top = w.topk(2, dim=-1).values
        self.pre[site] = (top[..., 0] - top[..., 1]).float()
        return y


def _relu(x, dec):
    return dec.relu(x) if dec is not None else F.relu(x)


def _pool(x, dec):
    return dec.max_pool(x) if dec is not None else F.max_pool2d(x, 2, 2)


# --------------------------------------------------------------------------- front end
def vgg_frontend(spec: torch.Tensor, P: Dict[str, torch.Tensor], dec: Optional[Decisions] = None) -> torch.Tensor:
    """models/asr/transformer.py:42-53 -- 2x(conv3x3+ReLU), pool, 2x(conv3x3+ReLU), pool."""
    h = _relu(F.conv2d(spec, P["conv.0.weight"], P["conv.0.bias"], padding=1), dec)
    h = _relu(F.conv2d(h, P["conv.2.weight"], P["conv.2.bias"], padding=1), dec)
    h = _pool(h, dec)
    h = _relu(F.conv2d(h, P["conv.5.weight"], P["conv.5.bias"], padding=1), dec)
    h = _relu(F.conv2d(h, P["conv.7.weight"], P["conv.7.bias"], padding=1), dec)
    return _pool(h, dec)


def emb_frontend(spec: torch.Tensor, P: Dict[str, torch.Tensor], eps: float = 1e-5, training: bool = True,
                 state: Optional[Dict[str, torch.Tensor]] = None, momentum: float = 0.1, dec: Optional[Decisions] = None) -> torch.Tensor:
    """models/asr/transformer.py:33-40 -- strided conv + BatchNorm2d + Hardtanh(0, 20), twice.
    nn.BatchNorm2d semantics: training -> batch statistics, and `state` (conv.{1,4}.running_mean / running_var /
    num_batches_tracked, the module's buffers) is updated in place with momentum 0.1 and the unbiased variance;
    eval -> `state` normalises (the reference validates with model.eval(), trainer/asr/trainer.py:123)."""
    def bn(h, i):
        rm = state[f"conv.{i}.running_mean"] if state is not None else None
        rv = state[f"conv.{i}.running_var"] if state is not None else None
        if training and state is not None:
            state[f"conv.{i}.num_batches_tracked"] += 1
        return F.batch_norm(h, rm, rv, P[f"conv.{i}.weight"], P[f"conv.{i}.bias"], training=training, momentum=momentum, eps=eps)

    h = F.conv2d(spec, P["conv.0.weight"], P["conv.0.bias"], stride=(2, 2), padding=(0, 10))
    clamp = (lambda t: dec.hardtanh(t)) if dec is not None else (lambda t: torch.clamp(t, 0.0, 20.0))
    h = clamp(bn(h, 1))
    h = F.conv2d(h, P["conv.3.weight"], P["conv.3.bias"], stride=(2, 1))
    return clamp(bn(h, 4))


def bn_initial_state(dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Fresh nn.BatchNorm2d(32) buffers of the emb_cnn front end (running_mean 0, running_var 1, num_batches_tracked 0)."""
    st = {}
    for i in (1, 4):
        st[f"conv.{i}.running_mean"] = torch.zeros(32, dtype=dtype)
        st[f"conv.{i}.running_var"] = torch.ones(32, dtype=dtype)
        st[f"conv.{i}.num_batches_tracked"] = torch.zeros((), dtype=torch.long)
    return st


def flatten_features(h: torch.Tensor) -> torch.Tensor:
    """models/asr/transformer.py:74-76 -- (B,C,F,T) -> (B,T,C*F), feature index c*F+f."""
    b, c, f, t = h.shape
    return h.reshape(b, c * f, t).transpose(1, 2).contiguous()


# --------------------------------------------------------------------------- masks
def length_non_pad(n_batch: int, t: int, lengths) -> torch.Tensor:
    """models/common_layers.py:33-38 -- ones, zeroed from lengths[i] on (B x T)."""
    m = torch.ones(n_batch, t)
    for i in range(n_batch):
        m[i, int(lengths[i]):] = 0
    return m


def length_key_mask(n_batch: int, t_k: int, lengths, t_q: int) -> torch.Tensor:
    """models/common_layers.py:57-64 -- True where the key frame is padding (B x Tq x Tk)."""
    pad = length_non_pad(n_batch, t_k, lengths).lt(1)
    return pad.unsqueeze(1).expand(-1, t_q, -1)


# --------------------------------------------------------------------------- blocks
def scaled_dot_attention(q, k, v, mask, temperature: float):
    """models/common_layers.py:215-223 (dropout omitted: oracle runs with p=0)."""
    s = torch.bmm(q, k.transpose(1, 2)) / temperature
    if mask is not None:
        s = s.masked_fill(mask, float("-inf"))
    p = torch.softmax(s, dim=2)
    return torch.bmm(p, v), p


def multi_head_attention(xq, xk, xv, mask, P, pre: str, H: int, dk: int, dv: int):
    """models/common_layers.py:170-200 -- head-major (h*B+b) batching, post-LN residual."""
    B, Tq, _ = xq.shape
    Tk = xk.shape[1]
    q = F.linear(xq, P[pre + "query_linear.weight"], P[pre + "query_linear.bias"]).view(B, Tq, H, dk)
    k = F.linear(xk, P[pre + "key_linear.weight"], P[pre + "key_linear.bias"]).view(B, Tk, H, dk)
    v = F.linear(xv, P[pre + "value_linear.weight"], P[pre + "value_linear.bias"]).view(B, Tk, H, dv)
    q = q.permute(2, 0, 1, 3).reshape(H * B, Tq, dk)
    k = k.permute(2, 0, 1, 3).reshape(H * B, Tk, dk)
    v = v.permute(2, 0, 1, 3).reshape(H * B, Tk, dv)
    m = mask.repeat(H, 1, 1) if mask is not None else None
    o, _ = scaled_dot_attention(q, k, v, m, float(dk) ** 0.5)
    o = o.view(H, B, Tq, dv).permute(1, 2, 0, 3).reshape(B, Tq, H * dv)
    o = F.linear(o, P[pre + "output_linear.weight"], P[pre + "output_linear.bias"])
    d = xq.shape[-1]
    return F.layer_norm(o + xq, (d,), P[pre + "layer_norm.weight"], P[pre + "layer_norm.bias"])


def conv_ffn(x, P, pre: str, dec: Optional[Decisions] = None):
    """models/common_layers.py:135-142 -- Conv1d(k=1) == per-token linear; weights (out,in,1)."""
    w1 = P[pre + "conv_1.weight"].squeeze(-1)
    w2 = P[pre + "conv_2.weight"].squeeze(-1)
    h = _relu(F.linear(x, w1, P[pre + "conv_1.bias"]), dec)
    y = F.linear(h, w2, P[pre + "conv_2.bias"])
    d = x.shape[-1]
    return F.layer_norm(y + x, (d,), P[pre + "layer_norm.weight"], P[pre + "layer_norm.bias"])


def encoder_forward(feats, lengths, P, cfg: OracleConfig, dec: Optional[Decisions] = None):
    """models/asr/transformer.py:157-180 and :195-203.  Note quirk Q1: `lengths` are raw
    frame counts compared against the (possibly down-sampled) feature length."""
    B, T, _ = feats.shape
    keep = length_non_pad(B, T, lengths).unsqueeze(-1).to(feats.dtype)
    mask = length_key_mask(B, T, lengths, T)
    d = cfg.dim_model
    x = F.linear(feats, P["encoder.input_linear.weight"], P["encoder.input_linear.bias"])
    x = F.layer_norm(x, (d,), P["encoder.layer_norm_input.weight"], P["encoder.layer_norm_input.bias"])
    x = x + sinusoid_table(T, d, feats.dtype).unsqueeze(0)
    for l in range(cfg.num_layers):
        pre = f"encoder.layers.{l}."
        x = multi_head_attention(x, x, x, mask, P, pre + "self_attn.", cfg.num_heads, cfg.dim_key, cfg.dim_value)
        x = x * keep
        x = conv_ffn(x, P, pre + "pos_ffn.", dec)
        x = x * keep
    return x


def preprocess_targets(padded_target: torch.Tensor, tgt_max_len: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """models/asr/transformer.py:254-266 + common_layers.py:14-22.
    seq_in = [SOS, y...] padded with EOS; seq_out = [y..., EOS] padded with PAD; both tgt_max_len."""
    B = padded_target.shape[0]
    seq_in = torch.full((B, tgt_max_len), EOS, dtype=torch.long)
    seq_out = torch.full((B, tgt_max_len), PAD, dtype=torch.long)
    for i in range(B):
        y = padded_target[i][padded_target[i] != PAD]
        n = y.numel()
        seq_in[i, 0] = SOS
        seq_in[i, 1:n + 1] = y
        seq_out[i, :n] = y
        seq_out[i, n] = EOS
    return seq_in, seq_out


def decoder_forward(padded_target, enc_out, enc_lengths, P, cfg: OracleConfig, dec: Optional[Decisions] = None):
    """models/asr/transformer.py:268-305 and :533-545."""
    seq_in, seq_out = preprocess_targets(padded_target, cfg.tgt_max_len)
    B, Tt = seq_in.shape
    Te = enc_out.shape[1]
    d = cfg.dim_model
    keep = seq_in.ne(EOS).to(enc_out.dtype).unsqueeze(-1)                      # :282
    causal = torch.triu(torch.ones(Tt, Tt, dtype=torch.uint8), diagonal=1)    # common_layers.py:66-74
    keypad = seq_in.eq(EOS).unsqueeze(1).expand(-1, Tt, -1)                   # common_layers.py:46-55
    self_mask = (keypad.to(torch.uint8) + causal.unsqueeze(0)).gt(0)          # :286
    cross_mask = length_key_mask(B, Te, enc_lengths, Tt)                      # :288-290
    scale = d ** -0.5 if cfg.emb_trg_sharing else 1.0                         # :248-252
    x = F.embedding(seq_in, P["decoder.trg_embedding.weight"], padding_idx=PAD) * scale
    x = x + sinusoid_table(Tt, d, enc_out.dtype).unsqueeze(0)                 # :292-293
    for l in range(cfg.num_layers):
        pre = f"decoder.layers.{l}."
        x = multi_head_attention(x, x, x, self_mask, P, pre + "self_attn.", cfg.num_heads, cfg.dim_key, cfg.dim_value)
        x = x * keep
        x = multi_head_attention(x, enc_out, enc_out, cross_mask, P, pre + "encoder_attn.", cfg.num_heads,
                                 cfg.dim_key, cfg.dim_value)
        x = x * keep
        x = conv_ffn(x, P, pre + "pos_ffn.", dec)
        x = x * keep
    logits = F.linear(x, P["decoder.output_linear.weight"])                   # :302 (no bias)
    return logits, seq_out


def transformer_forward(P, cfg: OracleConfig, spec, lengths, padded_target, training: bool = True, bn_state=None,
                        dec: Optional[Decisions] = None):
    """models/asr/transformer.py:59-85 -> (pred, gold, hyp_seq).  training / bn_state: BatchNorm mode of the emb_cnn front end;
    dec: record or replay the discrete decisions (class Decisions).  Site order: front end, encoder FFNs, decoder FFNs."""
    if cfg.feat_extractor == "vgg_cnn":
        h = flatten_features(vgg_frontend(spec, P, dec))
    elif cfg.feat_extractor == "emb_cnn":
        h = flatten_features(emb_frontend(spec, P, training=training, state=bn_state, dec=dec))
    else:
        h = flatten_features(spec)
    enc = encoder_forward(h, lengths, P, cfg, dec)
    pred, gold = decoder_forward(padded_target, enc, lengths, P, cfg, dec)
    hyp = pred.argmax(dim=2)                                                   # :80-82 (topk k=1)
    return pred, gold, hyp


def greedy_decode(P, cfg: OracleConfig, enc_out: torch.Tensor, steps: int):
    """models/asr/transformer.py:316-394 (greedy_search without LM rescoring): every step re-decodes the whole prefix with
    a causal self-attention mask, NO cross-attention mask (:347) and an all-ones non-pad mask (:336), then takes the argmax
    of the last position (:375).  The reference always runs 300 steps and cuts at EOS when building strings; here the
    token ids [B, steps] are returned (and the top-2 logit margin of every step, for tie-aware comparisons)."""
    B = enc_out.shape[0]
    d = cfg.dim_model
    scale = d ** -0.5 if cfg.emb_trg_sharing else 1.0
    ys = torch.full((B, 1), SOS, dtype=torch.long)
    margins = []
    for _ in range(steps):
        t = ys.shape[1]
        causal = torch.triu(torch.ones(t, t, dtype=torch.uint8), diagonal=1).bool().unsqueeze(0).expand(B, -1, -1)
        x = F.embedding(ys, P["decoder.trg_embedding.weight"], padding_idx=PAD) * scale + sinusoid_table(t, d, enc_out.dtype).unsqueeze(0)
        for l in range(cfg.num_layers):
            pre = f"decoder.layers.{l}."
            x = multi_head_attention(x, x, x, causal, P, pre + "self_attn.", cfg.num_heads, cfg.dim_key, cfg.dim_value)
            x = multi_head_attention(x, enc_out, enc_out, None, P, pre + "encoder_attn.", cfg.num_heads, cfg.dim_key, cfg.dim_value)
            x = conv_ffn(x, P, pre + "pos_ffn.")
        logits = F.linear(x[:, -1], P["decoder.output_linear.weight"])
        top2 = logits.topk(2, dim=1).values
        margins.append(top2[:, 0] - top2[:, 1])
        ys = torch.cat([ys, logits.argmax(dim=1, keepdim=True)], dim=1)
    return ys[:, 1:], torch.stack(margins, dim=1)


# --------------------------------------------------------------------------- loss
def cross_entropy_loss(pred: torch.Tensor, gold: torch.Tensor, smoothing: float):
    """utils/metrics.py:115-132.  Smoothing uses eps/V off-target (weights sum to 1-eps/V, quirk Q7);
    PAD rows are dropped and the sum is divided by the number of non-PAD tokens."""
    V = pred.shape[-1]
    logits = pred.reshape(-1, V)
    y = gold.reshape(-1)
    valid = y.ne(PAD)
    n_word = int(valid.sum())
    logp = F.log_softmax(logits, dim=1)
    if smoothing > 0.0:
        w = torch.full_like(logits, smoothing / V)
        w.scatter_(1, (y * valid.long()).view(-1, 1), 1.0 - smoothing)
        per_row = -(w * logp).sum(dim=1)
    else:
        per_row = -logp.gather(1, y.view(-1, 1)).squeeze(1)
    return (per_row * valid.to(per_row.dtype)).sum() / n_word, n_word


def num_correct(pred: torch.Tensor, gold: torch.Tensor) -> int:
    """utils/metrics.py:89-94."""
    hyp = pred.reshape(-1, pred.shape[-1]).argmax(dim=1)
    y = gold.reshape(-1)
    return int((hyp.eq(y) & y.ne(PAD)).sum())


def noam_rate(step: int, model_size: int, factor: float, warmup: int, min_lr: float) -> float:
    """utils/optimizer.py:27-32 (model_size is args.dim_input, quirk Q6)."""
    return max(min_lr, factor * (model_size ** -0.5 * min(step ** -0.5, step * warmup ** -1.5)))


def adam_reference(p, g, m, v, step: int, lr: float, b1=0.9, b2=0.98, eps=1e-9):
    """torch.optim.Adam as configured at utils/functions.py:107 (no weight decay, no amsgrad)."""
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    mhat = m / (1 - b1 ** step)
    vhat = v / (1 - b2 ** step)
    return p - lr * mhat / (vhat.sqrt() + eps), m, v


# --------------------------------------------------------------------------- helpers
def init_params(cfg: OracleConfig, seed: int = 123456, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Parameter set with the reference's names/shapes and its init rule:
    every tensor with dim>1 is xavier_uniform (transformer.py:55-57), biases keep the
    nn.Linear/Conv default, LayerNorm/BatchNorm affine = (1, 0)."""
    g = torch.Generator().manual_seed(seed)
    P: Dict[str, torch.Tensor] = {}

    def xavier(*shape):
        fan_out = shape[0] * (math.prod(shape[2:]) if len(shape) > 2 else 1)
        fan_in = shape[1] * (math.prod(shape[2:]) if len(shape) > 2 else 1)
        a = math.sqrt(6.0 / (fan_in + fan_out))
        return (torch.rand(*shape, generator=g, dtype=torch.float64) * 2 - 1).mul_(a).to(dtype)

    def bias(n, fan_in):
        a = 1.0 / math.sqrt(fan_in)
        return (torch.rand(n, generator=g, dtype=torch.float64) * 2 - 1).mul_(a).to(dtype)

    def lin(name, out_f, in_f, with_bias=True):
        P[name + ".weight"] = xavier(out_f, in_f)
        if with_bias:
            P[name + ".bias"] = bias(out_f, in_f)

    def ln(name, n):
        P[name + ".weight"] = torch.ones(n, dtype=dtype)
        P[name + ".bias"] = torch.zeros(n, dtype=dtype)

    def conv(name, co, ci, kh, kw):
        P[name + ".weight"] = xavier(co, ci, kh, kw)
        P[name + ".bias"] = bias(co, ci * kh * kw)

    if cfg.feat_extractor == "vgg_cnn":
        conv("conv.0", 64, 1, 3, 3); conv("conv.2", 64, 64, 3, 3)
        conv("conv.5", 128, 64, 3, 3); conv("conv.7", 128, 128, 3, 3)
    elif cfg.feat_extractor == "emb_cnn":
        conv("conv.0", 32, 1, 41, 11); ln("conv.1", 32)
        conv("conv.3", 32, 32, 21, 11); ln("conv.4", 32)
    d, H, dk, dv, di = cfg.dim_model, cfg.num_heads, cfg.dim_key, cfg.dim_value, cfg.dim_inner

    def mha(pre):
        lin(pre + "query_linear", H * dk, d); lin(pre + "key_linear", H * dk, d)
        lin(pre + "value_linear", H * dv, d); ln(pre + "layer_norm", d)
        lin(pre + "output_linear", d, H * dv)

    def ffn(pre):
        P[pre + "conv_1.weight"] = xavier(di, d, 1); P[pre + "conv_1.bias"] = bias(di, d)
        P[pre + "conv_2.weight"] = xavier(d, di, 1); P[pre + "conv_2.bias"] = bias(d, di)
        ln(pre + "layer_norm", d)

    lin("encoder.input_linear", d, cfg.dim_input); ln("encoder.layer_norm_input", d)
    for l in range(cfg.num_layers):
        mha(f"encoder.layers.{l}.self_attn."); ffn(f"encoder.layers.{l}.pos_ffn.")
    P["decoder.trg_embedding.weight"] = xavier(cfg.vocab, d)
    for l in range(cfg.num_layers):
        mha(f"decoder.layers.{l}.self_attn."); mha(f"decoder.layers.{l}.encoder_attn.")
        ffn(f"decoder.layers.{l}.pos_ffn.")
    if cfg.emb_trg_sharing:
        P["decoder.output_linear.weight"] = P["decoder.trg_embedding.weight"]
    else:
        P["decoder.output_linear.weight"] = xavier(cfg.vocab, d)
    return P


def synthetic_batch(cfg: OracleConfig, batch: int, t_src: int, seed: int = 0, ragged: bool = True):
    """Synthetic utterances shaped like utils/data_loader.py:182-214 (_collate_fn): zero-padded
    spectrograms sorted by length (descending), 0-padded int64 targets."""
    g = torch.Generator().manual_seed(seed)
    spec = torch.randn(batch, 1, cfg.freq, t_src, generator=g)
    if ragged:
        lens = (t_src * (0.5 + 0.5 * torch.rand(batch, generator=g))).long().clamp(1, t_src)
        lens[0] = t_src
        lens, _ = torch.sort(lens, descending=True)
    else:
        lens = torch.full((batch,), t_src, dtype=torch.long)
    for i in range(batch):
        spec[i, :, :, int(lens[i]):] = 0
    L = cfg.tgt_max_len - 1
    tgt = torch.randint(3, cfg.vocab, (batch, L), generator=g)
    if ragged:
        tl = (L * (0.4 + 0.6 * torch.rand(batch, generator=g))).long().clamp(1, L)
        tl[0] = L
        for i in range(batch):
            tgt[i, int(tl[i]):] = PAD
    return spec, lens.to(torch.int32), tgt


def forward_backward(P, cfg: OracleConfig, spec, lengths, tgt, smoothing: float, dec: Optional[Decisions] = None):
    """One oracle training step (no optimizer): returns pred, gold, hyp, loss and grads by name."""
    Pg = {k: v.detach().clone().requires_grad_(True) for k, v in P.items()}
    if cfg.emb_trg_sharing:
        Pg["decoder.output_linear.weight"] = Pg["decoder.trg_embedding.weight"]
    pred, gold, hyp = transformer_forward(Pg, cfg, spec, lengths, tgt, dec=dec)
    loss, n_word = cross_entropy_loss(pred, gold, smoothing)
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in Pg.items()}
    return pred.detach(), gold, hyp, loss.detach(), n_word, grads
