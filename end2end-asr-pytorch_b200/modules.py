"""Host-side mirror of the reference's model interface for the hot path.

Same class names, constructor arguments, parameter names/shapes (state_dict compatible) and forward
signatures as /root/reference/models/common_layers.py and models/asr/transformer.py -- but every forward is
a composition of libb200asr kernels (ops.py).  The forward bodies are written against attribute names that
exist in BOTH this mirror and the reference classes, so ``install()`` can bind the very same functions onto
the reference's classes (class-level patch).

Not mirrored (out of scope): beam_search, LM rescoring, the Linear-variant FFN (dead code).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import ops

PAD_TOKEN, SOS_TOKEN, EOS_TOKEN = 0, 1, 2     # utils/constant.py:102-104


# ------------------------------------------------------------------------------------------------ helpers
def _drop_p(module: nn.Module, dropout: nn.Dropout) -> float:
    return float(dropout.p) if module.training else 0.0


def _as_u8_mask(mask, B, Tq, Tk):
    """Reference-style bool/uint8 (B,Tq,Tk) mask -> contiguous uint8 (compat path only; the model-level forwards
    never build dense masks)."""
    if mask is None:
        return None
    m = mask.expand(B, Tq, Tk)
    return m.to(torch.uint8).contiguous()


def _mha_core(self, query, key, value, key_pad=None, dense_mask=None, causal=False, row_scale=None):
    """MultiHeadAttention body (models/common_layers.py:170-200) on fused kernels.

    Q/K/V projections write token-major [B,T,H*d]; the attention kernel addresses heads through strides, so
    the permute/contiguous copies of :185-187 and :194-195 do not exist.  `row_scale` (the non-pad mask that the
    layer applies right after, transformer.py:198/536/540) is folded into the LayerNorm epilogue."""
    B, Tq, _ = query.shape
    Tk = key.shape[1]
    H, dk, dv = self.num_heads, self.dim_key, self.dim_value
    p_attn = _drop_p(self, self.attention.dropout)
    scale = 1.0 / float(self.attention.temperature)
    ql, kl, vl = self.query_linear, self.key_linear, self.value_linear
    link = None
    if key is value:
        link = ops.ResidualLink() if torch.is_grad_enabled() else None
        # projections + attention as one autograd node: Q/K/V (or K/V) run as ONE GEMM when their weights lie back to back
        # in memory (optim.FlatParams), and the attention kernels address the packed output through strides
        o = ops.AttnProjFn.apply(query, key, ql.weight, ql.bias, kl.weight, kl.bias, vl.weight, vl.bias, H, dk, dv,
                                 key_pad, dense_mask, causal, scale, p_attn, link)          # [B,Tq,H*dv], token-major
    else:
        q = ops.LinearFn.apply(query, ql.weight, ql.bias).view(B, Tq, H, dk).permute(0, 2, 1, 3)
        k = ops.LinearFn.apply(key, kl.weight, kl.bias).view(B, Tk, H, dk).permute(0, 2, 1, 3)
        v = ops.LinearFn.apply(value, vl.weight, vl.bias).view(B, Tk, H, dv).permute(0, 2, 1, 3)
        o = ops.SdpaFn.apply(q, k, v, key_pad, dense_mask, causal, scale, p_attn)         # (B,H,Tq,dv) view of [B,Tq,H,dv]
        o = o.permute(0, 2, 1, 3).reshape(B, Tq, H * dv)                                  # free: memory is token-major
    o = ops.LinearFn.apply(o, self.output_linear.weight, self.output_linear.bias)
    ln = self.layer_norm
    return ops.AddLNFn.apply(o, query, ln.weight, ln.bias, None, row_scale, ln.eps, _drop_p(self, self.dropout), link)


def _ffn_core(self, x, row_scale=None):
    """PositionwiseFeedForwardWithConv body (models/common_layers.py:135-142)."""
    link = ops.ResidualLink() if torch.is_grad_enabled() else None
    y = ops.FFNFn.apply(x, self.conv_1.weight, self.conv_1.bias, self.conv_2.weight, self.conv_2.bias, link)
    ln = self.layer_norm
    return ops.AddLNFn.apply(y, x, ln.weight, ln.bias, None, row_scale, ln.eps, _drop_p(self, self.dropout), link)


# ------------------------------------------------------------------------------------------------ forwards (bindable)
def sdpa_forward(self, q, k, v, mask=None):
    """ScaledDotProductAttention.forward(q, k, v, mask) -- models/common_layers.py:211-225.
    q,k,v are head-major (H*B) x T x d exactly as the reference passes them and `mask` is the (H*B) x Tq x Tk
    tensor MultiHeadAttention built with mask.repeat(H,1,1); each (h,b) slice is treated as its own batch entry.
    Returns (output, None): attention probabilities are never materialised (every caller discards them)."""
    HB, Tq, dk = q.shape
    Tk, dv = k.shape[1], v.shape[2]
    dense = _as_u8_mask(mask, HB, Tq, Tk)
    as4 = lambda t: t.contiguous().unsqueeze(1)          # (HB,1,T,d): one "head" per batch entry
    p = float(self.dropout.p) if self.training else 0.0
    o = ops.SdpaFn.apply(as4(q), as4(k), as4(v), None, dense, False, 1.0 / float(self.temperature), p)
    return o.reshape(HB, Tq, dv), None


def mha_forward(self, query, key, value, mask=None):
    """MultiHeadAttention.forward(query, key, value, mask) -> (output, attn=None) -- common_layers.py:170-200."""
    B, Tq, _ = query.shape
    dense = _as_u8_mask(mask, B, Tq, key.shape[1])
    return _mha_core(self, query, key, value, dense_mask=dense), None


def ffn_forward(self, x):
    """PositionwiseFeedForwardWithConv.forward(x) -- common_layers.py:135-142."""
    return _ffn_core(self, x)


def encoder_layer_forward(self, enc_input, non_pad_mask=None, self_attn_mask=None):
    """EncoderLayer.forward -- models/asr/transformer.py:195-203 (dense-mask compatible signature)."""
    B, T, _ = enc_input.shape
    rs = non_pad_mask.reshape(-1) if non_pad_mask is not None else None
    x = _mha_core(self.self_attn, enc_input, enc_input, enc_input, dense_mask=_as_u8_mask(self_attn_mask, B, T, T), row_scale=rs)
    return _ffn_core(self.pos_ffn, x, row_scale=rs), None


def decoder_layer_forward(self, decoder_input, encoder_output, non_pad_mask=None, self_attn_mask=None, dec_enc_attn_mask=None):
    """DecoderLayer.forward -- models/asr/transformer.py:533-545 (dense-mask compatible signature)."""
    B, T, _ = decoder_input.shape
    Te = encoder_output.shape[1]
    rs = non_pad_mask.reshape(-1) if non_pad_mask is not None else None
    x = _mha_core(self.self_attn, decoder_input, decoder_input, decoder_input,
                  dense_mask=_as_u8_mask(self_attn_mask, B, T, T), row_scale=rs)
    x = _mha_core(self.encoder_attn, x, encoder_output, encoder_output,
                  dense_mask=_as_u8_mask(dec_enc_attn_mask, B, T, Te), row_scale=rs)
    return _ffn_core(self.pos_ffn, x, row_scale=rs), None, None


def _encoder_body(self, feats, input_lengths, cf_order=None):
    """Encoder.forward body (models/asr/transformer.py:157-180).  `feats` is [B,T,D]; when cf_order=(C,F) its
    feature axis is ordered f*C+c (channels-last conv output) and input_linear's columns are permuted to match.
    Masks are a key-valid byte vector per utterance built from the RAW lengths (quirk Q1 preserved)."""
    B, T, D = feats.shape
    key_pad, non_pad = ops.length_masks(input_lengths, B, T, feats.device)
    rs = non_pad.view(-1)
    w = self.input_linear.weight
    if cf_order is not None:
        w = ops.PermuteColsFn.apply(w, cf_order[0], cf_order[1])
    x = ops.LinearFn.apply(feats, w, self.input_linear.bias)
    ln = self.layer_norm_input
    pe = self.positional_encoding.pe[0, :T]
    x = ops.AddLNFn.apply(x, None, ln.weight, ln.bias, pe, None, ln.eps, 0.0)
    for layer in self.layers:
        x = _mha_core(layer.self_attn, x, x, x, key_pad=key_pad, row_scale=rs)
        x = _ffn_core(layer.pos_ffn, x, row_scale=rs)
    return x


def encoder_forward(self, padded_input, input_lengths):
    """Encoder.forward(padded_input, input_lengths) -> (output, [None]*L)."""
    return _encoder_body(self, padded_input, input_lengths), [None] * len(self.layers)


def decoder_forward(self, padded_input, encoder_padded_outputs, encoder_input_lengths):
    """Decoder.forward -- models/asr/transformer.py:268-305: (pred, gold, [None]*L, [None]*L).
    Target preprocessing, masks, embedding+PE run on the device without host syncs."""
    Tt = int(self.trg_max_length)
    seq_in, seq_out, key_pad, non_pad = ops.preprocess_targets(padded_input, Tt)
    B = seq_in.shape[0]
    Te = encoder_padded_outputs.shape[1]
    enc_key_pad, _ = ops.length_masks(encoder_input_lengths, B, Te, encoder_padded_outputs.device)
    rs = non_pad.view(-1)
    pe = self.positional_encoding.pe[0, :Tt]
    x = ops.EmbedFn.apply(seq_in, self.trg_embedding.weight, pe, float(self.x_logit_scale), _drop_p(self, self.dropout), PAD_TOKEN)
    for layer in self.layers:
        x = _mha_core(layer.self_attn, x, x, x, key_pad=key_pad, causal=True, row_scale=rs)
        x = _mha_core(layer.encoder_attn, x, encoder_padded_outputs, encoder_padded_outputs, key_pad=enc_key_pad, row_scale=rs)
        x = _ffn_core(layer.pos_ffn, x, row_scale=rs)
    pred = ops.LinearFn.apply(x, self.output_linear.weight, None)
    n = len(self.layers)
    return pred, seq_out, [None] * n, [None] * n


def greedy_decode_ids(self, encoder_padded_outputs, steps=300):
    """Token ids of Decoder.greedy_search (models/asr/transformer.py:316-394, no LM rescoring) on the fused kernels:
    full-prefix re-decode per step exactly as the reference (causal self-attention, no cross-attention mask, all-ones
    non-pad mask, argmax of the last position).  Returns int64 [B, steps]; string building / EOS cut stay with the caller.
    (The KV-cached incremental decoder is the 'next' row 3 of SURVEY.md §8f.)"""
    B = encoder_padded_outputs.shape[0]
    dev = encoder_padded_outputs.device
    was_training = self.training
    self.eval()
    try:
        with torch.no_grad():
            ys = torch.full((B, 1), SOS_TOKEN, dtype=torch.long, device=dev)
            for _ in range(steps):
                t = ys.shape[1]
                pe = self.positional_encoding.pe[0, :t]
                x = ops.EmbedFn.apply(ys, self.trg_embedding.weight, pe, float(self.x_logit_scale), 0.0, PAD_TOKEN)
                for layer in self.layers:
                    x = _mha_core(layer.self_attn, x, x, x, causal=True)
                    x = _mha_core(layer.encoder_attn, x, encoder_padded_outputs, encoder_padded_outputs)
                    x = _ffn_core(layer.pos_ffn, x)
                last = x[:, -1].contiguous()
                logits = ops.LinearFn.apply(last, self.output_linear.weight, None)
                ys = torch.cat([ys, ops.argmax_rows(logits).view(B, 1)], dim=1)
    finally:
        self.train(was_training)
    return ys[:, 1:]


def greedy_decode_cached(self, encoder_padded_outputs, steps=300, stop_at_eos=False, check_every=8):
    """Incremental greedy decode with self- and cross-attention K/V caches: the same token ids as
    greedy_decode_ids / the reference's greedy_search, at O(T) instead of O(T^2) work -- per step only the new position is
    projected, its K/V rows are appended to per-layer caches, and single-query attention runs over the cache (self) and
    over the once-projected encoder K/V (cross, no mask as in transformer.py:347).

    stop_at_eos: the EOS cut of transformer.py:385-393 stays on the device -- ids from an utterance's first EOS on are -1
    (the reference drops them when it builds the strings), a device counter tracks finished utterances and the loop ends
    as soon as every utterance has emitted EOS (checked every `check_every` steps: one 4-byte read, no id traffic)
    instead of always running the reference's 300 fixed steps."""
    B, Te, _ = encoder_padded_outputs.shape
    dev = encoder_padded_outputs.device
    was_training = self.training
    self.eval()
    try:
        with torch.no_grad():
            H = self.layers[0].self_attn.num_heads
            dk, dv = self.layers[0].self_attn.dim_key, self.layers[0].self_attn.dim_value
            cross, kc, vc = [], [], []
            for layer in self.layers:
                ca = layer.encoder_attn
                k = ops.LinearFn.apply(encoder_padded_outputs, ca.key_linear.weight, ca.key_linear.bias).view(B, Te, H, dk).permute(0, 2, 1, 3)
                v = ops.LinearFn.apply(encoder_padded_outputs, ca.value_linear.weight, ca.value_linear.bias).view(B, Te, H, dv).permute(0, 2, 1, 3)
                cross.append((k, v))
                kc.append(torch.empty((B, steps, H * dk), device=dev, dtype=torch.float32))
                vc.append(torch.empty((B, steps, H * dv), device=dev, dtype=torch.float32))
            ys = torch.full((B, steps), -1, dtype=torch.long, device=dev) if stop_at_eos else torch.empty((B, steps), dtype=torch.long, device=dev)
            tok = torch.full((B, 1), SOS_TOKEN, dtype=torch.long, device=dev)
            finished = torch.zeros(B + 1, dtype=torch.int32, device=dev) if stop_at_eos else None      # [B] flags + [1] counter

            def attend(mod, x, k4, v4):
                q = ops.LinearFn.apply(x, mod.query_linear.weight, mod.query_linear.bias).view(B, 1, H, dk).permute(0, 2, 1, 3)
                o = ops.SdpaFn.apply(q, k4, v4, None, None, False, 1.0 / float(mod.attention.temperature), 0.0)
                o = ops.LinearFn.apply(o.permute(0, 2, 1, 3).reshape(B, 1, H * dv), mod.output_linear.weight, mod.output_linear.bias)
                return ops.AddLNFn.apply(o, x, mod.layer_norm.weight, mod.layer_norm.bias, None, None, mod.layer_norm.eps, 0.0)

            for t in range(steps):
                pe = self.positional_encoding.pe[0, t:t + 1]
                x = ops.EmbedFn.apply(tok, self.trg_embedding.weight, pe, float(self.x_logit_scale), 0.0, PAD_TOKEN)
                for li, layer in enumerate(self.layers):
                    sa = layer.self_attn
                    kc[li][:, t:t + 1] = ops.LinearFn.apply(x, sa.key_linear.weight, sa.key_linear.bias)
                    vc[li][:, t:t + 1] = ops.LinearFn.apply(x, sa.value_linear.weight, sa.value_linear.bias)
                    k4 = kc[li][:, :t + 1].view(B, t + 1, H, dk).permute(0, 2, 1, 3)
                    v4 = vc[li][:, :t + 1].view(B, t + 1, H, dv).permute(0, 2, 1, 3)
                    x = attend(sa, x, k4, v4)
                    x = attend(layer.encoder_attn, x, cross[li][0], cross[li][1])
                    x = _ffn_core(layer.pos_ffn, x)
                logits = ops.LinearFn.apply(x.view(B, -1), self.output_linear.weight, None)
                if stop_at_eos:
                    tok = ops.greedy_step(logits, ys, finished, t)
                    if (t + 1) % check_every == 0 and int(finished[B].item()) == B:
                        break
                else:
                    tok = ops.argmax_rows(logits).view(B, 1)
                    ys[:, t:t + 1] = tok
    finally:
        self.train(was_training)
    return ys


def _front_end(self, padded_input):
    """CNN front end + flatten of Transformer.forward (models/asr/transformer.py:70-76).
    Returns (feats [B,T',D], cf_order or None)."""
    if self.feat_extractor == "vgg_cnn":
        c = self.conv
        h = ops.VggFrontendFn.apply(padded_input, c[0].weight, c[0].bias, c[2].weight, c[2].bias, c[5].weight, c[5].bias,
                                    c[7].weight, c[7].bias)
        B, T4, F4, C = h.shape
        return h.view(B, T4, F4 * C), (C, F4)
    if self.feat_extractor == "emb_cnn":
        c = self.conv
        # nn.BatchNorm2d state (running_mean, running_var, num_batches_tracked, momentum): updated by the kernel in training
        # mode, used for the normalisation in eval mode -- exactly what the stock module does (transformer.py:34,38)
        def bn_state(m):
            if not m.track_running_stats or m.running_mean is None:
                return None
            return (m.running_mean, m.running_var, m.num_batches_tracked, m.momentum)
        training = self.training or bn_state(c[1]) is None or bn_state(c[4]) is None
        h = ops.EmbFrontendFn.apply(padded_input, c[0].weight, c[0].bias, c[1].weight, c[1].bias, c[3].weight, c[3].bias,
                                    c[4].weight, c[4].bias, float(c[1].eps), training, bn_state(c[1]), bn_state(c[4]))
        return h, None
    return ops.FlattenFn.apply(padded_input), None


def transformer_forward(self, padded_input, input_lengths, padded_target, verbose=False):
    """Transformer.forward -> (pred, gold, hyp_seq, gold_seq) -- models/asr/transformer.py:59-85."""
    ops._need_cuda(padded_input, padded_target)
    feats, cf = _front_end(self, padded_input)
    enc = _encoder_body(self.encoder, feats, input_lengths, cf)
    pred, gold, *_ = decoder_forward(self.decoder, padded_target, enc, input_lengths)
    hyp_seq = ops.argmax_rows(pred)
    return pred, gold, hyp_seq, gold


# ------------------------------------------------------------------------------------------------ mirror classes
class PositionalEncoding(nn.Module):
    """models/common_layers.py:76-98."""

    def __init__(self, dim_model, max_length=2000):
        super().__init__()
        pe = torch.zeros(max_length, dim_model)
        position = torch.arange(0, max_length).unsqueeze(1).float()
        exp_term = torch.exp(torch.arange(0, dim_model, 2).float() * -(math.log(10000.0) / dim_model))
        pe[:, 0::2] = torch.sin(position * exp_term)
        pe[:, 1::2] = torch.cos(position * exp_term)
        self.register_buffer("pe", pe.unsqueeze(0))

    def forward(self, input):
        return self.pe[:, :input.size(1)]


class ScaledDotProductAttention(nn.Module):
    def __init__(self, temperature, attn_dropout=0.1):
        super().__init__()
        self.temperature = temperature
        self.dropout = nn.Dropout(attn_dropout)
        self.softmax = nn.Softmax(dim=2)

    forward = sdpa_forward


class MultiHeadAttention(nn.Module):
    def __init__(self, num_heads, dim_model, dim_key, dim_value, dropout=0.1):
        super().__init__()
        self.num_heads, self.dim_model, self.dim_key, self.dim_value = num_heads, dim_model, dim_key, dim_value
        self.query_linear = nn.Linear(dim_model, num_heads * dim_key)
        self.key_linear = nn.Linear(dim_model, num_heads * dim_key)
        self.value_linear = nn.Linear(dim_model, num_heads * dim_value)
        self.attention = ScaledDotProductAttention(temperature=float(dim_key) ** 0.5, attn_dropout=dropout)
        self.layer_norm = nn.LayerNorm(dim_model)
        self.output_linear = nn.Linear(num_heads * dim_value, dim_model)
        self.dropout = nn.Dropout(dropout)

    forward = mha_forward


class PositionwiseFeedForwardWithConv(nn.Module):
    def __init__(self, dim_model, dim_hidden, dropout=0.1):
        super().__init__()
        self.conv_1 = nn.Conv1d(dim_model, dim_hidden, 1)
        self.conv_2 = nn.Conv1d(dim_hidden, dim_model, 1)
        self.dropout = nn.Dropout(dropout)
        self.layer_norm = nn.LayerNorm(dim_model)

    forward = ffn_forward


class EncoderLayer(nn.Module):
    def __init__(self, num_heads, dim_model, dim_inner, dim_key, dim_value, dropout=0.1):
        super().__init__()
        self.self_attn = MultiHeadAttention(num_heads, dim_model, dim_key, dim_value, dropout=dropout)
        self.pos_ffn = PositionwiseFeedForwardWithConv(dim_model, dim_inner, dropout=dropout)

    forward = encoder_layer_forward


class Encoder(nn.Module):
    def __init__(self, num_layers, num_heads, dim_model, dim_key, dim_value, dim_input, dim_inner, dropout=0.1,
                 src_max_length=2500):
        super().__init__()
        self.dim_input, self.num_layers, self.num_heads = dim_input, num_layers, num_heads
        self.dim_model, self.dim_key, self.dim_value, self.dim_inner = dim_model, dim_key, dim_value, dim_inner
        self.src_max_length = src_max_length
        self.dropout = nn.Dropout(dropout)
        self.dropout_rate = dropout
        self.input_linear = nn.Linear(dim_input, dim_model)
        self.layer_norm_input = nn.LayerNorm(dim_model)
        self.positional_encoding = PositionalEncoding(dim_model, src_max_length)
        self.layers = nn.ModuleList([EncoderLayer(num_heads, dim_model, dim_inner, dim_key, dim_value, dropout=dropout)
                                     for _ in range(num_layers)])

    forward = encoder_forward


class DecoderLayer(nn.Module):
    def __init__(self, dim_model, dim_inner, num_heads, dim_key, dim_value, dropout=0.1):
        super().__init__()
        self.self_attn = MultiHeadAttention(num_heads, dim_model, dim_key, dim_value, dropout=dropout)
        self.encoder_attn = MultiHeadAttention(num_heads, dim_model, dim_key, dim_value, dropout=dropout)
        self.pos_ffn = PositionwiseFeedForwardWithConv(dim_model, dim_inner, dropout=dropout)

    forward = decoder_layer_forward


class Decoder(nn.Module):
    def __init__(self, id2label, num_src_vocab, num_trg_vocab, num_layers, num_heads, dim_emb, dim_model, dim_inner,
                 dim_key, dim_value, dropout=0.1, trg_max_length=1000, emb_trg_sharing=False):
        super().__init__()
        self.sos_id, self.eos_id = SOS_TOKEN, EOS_TOKEN
        self.id2label = id2label
        self.num_src_vocab, self.num_trg_vocab = num_src_vocab, num_trg_vocab
        self.num_layers, self.num_heads = num_layers, num_heads
        self.dim_emb, self.dim_model, self.dim_inner = dim_emb, dim_model, dim_inner
        self.dim_key, self.dim_value = dim_key, dim_value
        self.dropout_rate = dropout
        self.emb_trg_sharing = emb_trg_sharing
        self.trg_max_length = trg_max_length
        self.trg_embedding = nn.Embedding(num_trg_vocab, dim_emb, padding_idx=PAD_TOKEN)
        self.positional_encoding = PositionalEncoding(dim_model, max_length=trg_max_length)
        self.dropout = nn.Dropout(dropout)
        self.layers = nn.ModuleList([DecoderLayer(dim_model, dim_inner, num_heads, dim_key, dim_value, dropout=dropout)
                                     for _ in range(num_layers)])
        self.output_linear = nn.Linear(dim_model, num_trg_vocab, bias=False)
        if emb_trg_sharing:
            self.output_linear.weight = self.trg_embedding.weight
            self.x_logit_scale = dim_model ** -0.5
        else:
            self.x_logit_scale = 1.0

    forward = decoder_forward
    greedy_decode_ids = greedy_decode_ids
    greedy_decode_cached = greedy_decode_cached


class Transformer(nn.Module):
    """models/asr/transformer.py:16-85 (forward only; evaluate/beam search are out of scope)."""

    def __init__(self, encoder, decoder, feat_extractor="vgg_cnn"):
        super().__init__()
        self.encoder, self.decoder = encoder, decoder
        self.id2label = decoder.id2label
        self.feat_extractor = feat_extractor
        if feat_extractor == "emb_cnn":
            self.conv = nn.Sequential(
                nn.Conv2d(1, 32, kernel_size=(41, 11), stride=(2, 2), padding=(0, 10)), nn.BatchNorm2d(32),
                nn.Hardtanh(0, 20, inplace=True),
                nn.Conv2d(32, 32, kernel_size=(21, 11), stride=(2, 1)), nn.BatchNorm2d(32),
                nn.Hardtanh(0, 20, inplace=True))
        elif feat_extractor == "vgg_cnn":
            self.conv = nn.Sequential(
                nn.Conv2d(1, 64, 3, stride=1, padding=1), nn.ReLU(), nn.Conv2d(64, 64, 3, stride=1, padding=1), nn.ReLU(),
                nn.MaxPool2d(2, stride=2),
                nn.Conv2d(64, 128, 3, stride=1, padding=1), nn.ReLU(), nn.Conv2d(128, 128, 3, stride=1, padding=1), nn.ReLU(),
                nn.MaxPool2d(2, stride=2))
        for p in self.parameters():      # transformer.py:55-57 (quirk Q4)
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)

    forward = transformer_forward


def build_model(cfg, id2label=None) -> Transformer:
    """Equivalent of utils/functions.py:116-152 (init_transformer_model) from an ASRConfig."""
    id2label = id2label or {i: str(i) for i in range(cfg.vocab)}
    enc = Encoder(cfg.num_layers, num_heads=cfg.num_heads, dim_model=cfg.dim_model, dim_key=cfg.dim_key,
                  dim_value=cfg.dim_value, dim_input=cfg.dim_input, dim_inner=cfg.dim_inner,
                  src_max_length=cfg.src_max_len, dropout=cfg.dropout)
    dec = Decoder(id2label, num_src_vocab=cfg.vocab, num_trg_vocab=cfg.vocab, num_layers=cfg.num_layers,
                  num_heads=cfg.num_heads, dim_emb=cfg.dim_model, dim_model=cfg.dim_model, dim_inner=cfg.dim_inner,
                  dim_key=cfg.dim_key, dim_value=cfg.dim_value, trg_max_length=cfg.tgt_max_len, dropout=cfg.dropout,
                  emb_trg_sharing=cfg.emb_trg_sharing)
    return Transformer(enc, dec, feat_extractor=cfg.feat_extractor)
