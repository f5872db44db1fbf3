"""Model/shape configuration of the hot path (the subset of utils/constant.py flags the model reads)."""
from __future__ import annotations

from dataclasses import dataclass


@dataclass
class ASRConfig:
    num_layers: int = 4
    num_heads: int = 8
    dim_model: int = 512
    dim_key: int = 64
    dim_value: int = 64
    dim_inner: int = 2048
    vocab: int = 4364                 # 3 specials + data/labels/aishell_labels.json (train.py:47-57)
    feat_extractor: str = "vgg_cnn"   # 'vgg_cnn' | 'emb_cnn' | ''
    tgt_max_len: int = 100
    src_max_len: int = 4000
    freq: int = 161                   # sample_rate * window_size / 2 + 1
    dropout: float = 0.1
    label_smoothing: float = 0.1
    emb_trg_sharing: bool = False

    @property
    def dim_input(self) -> int:
        """utils/functions.py:120-130."""
        if self.feat_extractor == "vgg_cnn":
            return (self.freq // 2 // 2) * 128
        if self.feat_extractor == "emb_cnn":
            h = (self.freq - 41) // 2 + 1
            h = (h - 21) // 2 + 1
            return h * 32
        return self.freq

    def t_enc(self, t_src: int) -> int:
        """Encoder length for t_src input frames."""
        if self.feat_extractor == "vgg_cnn":
            return t_src // 2 // 2
        if self.feat_extractor == "emb_cnn":
            return ((t_src + 20 - 11) // 2 + 1) - 11 + 1
        return t_src


# BASELINE.json configs
BASELINE_CONFIGS = {
    "cfg1": dict(cfg=ASRConfig(num_layers=1, num_heads=2, dim_model=128, dim_key=64, dim_value=64, dim_inner=1024, vocab=32,
                               feat_extractor="", tgt_max_len=10), batch=2, t_src=50),
    "cfg2": dict(cfg=ASRConfig(), batch=32, t_src=800),
    "cfg3": dict(cfg=ASRConfig(num_layers=6, dim_key=32, dim_value=32, feat_extractor="emb_cnn"), batch=64, t_src=400),
    "cfg4": dict(cfg=ASRConfig(num_layers=6, tgt_max_len=150), batch=32, t_src=1000),       # per GPU; global 256 on 8
    "cfg5": dict(cfg=ASRConfig(num_layers=12, dim_model=768, dim_inner=3072, vocab=32, tgt_max_len=150), batch=16,
                 t_src=1600),                                                                # per GPU; global 128 on 8
}
