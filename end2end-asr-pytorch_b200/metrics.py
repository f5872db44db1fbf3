"""Mirror of utils/metrics.py:78-132 (calculate_metrics / calculate_loss, CE branches) on the fused CE kernel."""
from __future__ import annotations

import torch

from . import ops


def loss_and_stats(pred, gold, smoothing=0.0, reduction="mean"):
    """(loss, stats) without any host synchronisation.  stats = [sum, n_tokens, n_correct, mean, 1/n_tokens]."""
    return ops.CrossEntropyFn.apply(pred, gold, float(smoothing), reduction)


def calculate_loss(pred, gold, input_lengths=None, target_lengths=None, smoothing=0.0, loss_type="ce"):
    """utils/metrics.py:102-132.  Only the CE branches are on the hot path (CTC is vestigial)."""
    if loss_type != "ce":
        raise NotImplementedError("b200asr implements the 'ce' loss of the hot path only (ctc is out of scope)")
    loss, _ = loss_and_stats(pred, gold, smoothing)
    return loss


def calculate_metrics(pred, gold, input_lengths=None, target_lengths=None, smoothing=0.0, loss_type="ce"):
    """utils/metrics.py:78-95 -> (loss, num_correct).  num_correct is a Python int as in the reference (one D2H
    sync, the reference has the same at :94); use loss_and_stats() to stay asynchronous."""
    if loss_type != "ce":
        raise NotImplementedError("b200asr implements the 'ce' loss of the hot path only (ctc is out of scope)")
    loss, stats = loss_and_stats(pred, gold, smoothing)
    return loss, int(stats[2].item())
