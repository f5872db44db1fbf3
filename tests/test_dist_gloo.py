"""World-size-2 gloo test of the data-parallel host protocol (parallel.py): shard the batch, back-propagate the
UN-normalised loss sum, ONE all-reduce of [flat grads | sum-loss | n_tokens], divide by the GLOBAL token count --
must equal a single-rank step on the whole batch (the reference's DataParallel semantics).
The CUDA kernels cannot run here, so a small CPU model / loss / Adam stand in for them; the protocol code is the
product's."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import b200asr
from oracle import asr_oracle as O


class TinyModel(torch.nn.Module):
    """(src, lengths, tgt) -> (pred, gold, hyp, gold) with the hot path's signature."""

    def __init__(self, V=11):
        super().__init__()
        self.a = torch.nn.Linear(6, 16)
        self.b = torch.nn.Linear(16, V)

    def forward(self, src, lengths, tgt):
        pred = self.b(torch.tanh(self.a(src)))
        return pred, tgt, pred.argmax(-1), tgt


class _FrontEndBoundary(torch.autograd.Function):
    """Stands in for the CNN front end's backward entry (ops.VggFrontendFn.backward calls ops.frontend_backward_hook first)."""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        import importlib
        ops = importlib.import_module(b200asr.__name__ + ".ops")
        if ops.frontend_backward_hook is not None:
            ops.frontend_backward_hook()
        return g


class TinyModelWithFrontEnd(TinyModel):
    """A `conv` front end in front of the rest: its parameters go to the tail of the flat buffers and the head part of the
    gradient all-reduce is launched while its backward is still to run (DataParallelStep._reduce_head_early)."""

    def __init__(self, V=11):
        super().__init__(V)
        self.conv = torch.nn.Linear(6, 6)

    def forward(self, src, lengths, tgt):
        return super().forward(_FrontEndBoundary.apply(self.conv(src)), lengths, tgt)


def cpu_loss(pred, gold, smoothing, reduction="mean"):
    loss, n = O.cross_entropy_loss(pred, gold, smoothing)
    s = loss * n
    stats = torch.stack([s.detach(), torch.tensor(float(n)), torch.tensor(0.0), loss.detach(), torch.tensor(1.0 / n)])
    return (s if reduction == "sum" else loss), stats


class CpuAdam:
    def __init__(self, flat):
        self.flat, self.param_groups, self.t = flat, [dict(lr=0.0)], 0
        self.m, self.v = torch.zeros_like(flat.flat), torch.zeros_like(flat.flat)

    def zero_grad(self):
        self.flat.zero_grad()

    def step(self, grad_scale=1.0, grad_scale_dev=None):
        self.t += 1
        g = self.flat.flat_grad[:self.flat.numel] * grad_scale * (grad_scale_dev if grad_scale_dev is not None else 1.0)
        p, self.m, self.v = O.adam_reference(self.flat.flat, g, self.m, self.v, self.t, self.param_groups[0]["lr"])
        self.flat.flat.copy_(p)


def make_batch():
    g = torch.Generator().manual_seed(5)
    src = torch.randn(4, 7, 6, generator=g)
    tgt = torch.randint(1, 11, (4, 7), generator=g)
    tgt[0, 5:] = 0; tgt[2, 2:] = 0; tgt[3, 6:] = 0           # ranks see different token counts
    return src, torch.full((4,), 7), tgt


def run_steps(model, world_group, rank, world, steps=3):
    dp = b200asr.DataParallelStep(model, model_size=64, warmup=5, k_lr=1.0, min_lr=1e-6, smoothing=0.1, process_group=world_group,
                                  loss_fn=cpu_loss, adam_factory=CpuAdam)
    src, lens, tgt = make_batch()
    losses = []
    for _ in range(steps):
        s, l, t = b200asr.shard_batch(src, lens, tgt, rank, world) if world > 1 else (src, lens, tgt)
        dp.step(s, l, t)
        losses.append(float(dp.global_loss()))
    return losses, dp.flat.flat.clone()


def _worker(rank, world, port, q, cls="TinyModel"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(rank)          # ranks build DIFFERENT replicas: DataParallelStep broadcasts rank 0's (seed 0) parameters
    model = globals()[cls]()
    losses, flat = run_steps(model, None, rank, world)
    q.put((rank, losses, flat))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("cls", ["TinyModel", "TinyModelWithFrontEnd"])
def test_two_rank_step_equals_single_rank_global_batch(cls):
    torch.manual_seed(0)
    ref_model = globals()[cls]()
    ref_losses, ref_flat = run_steps(ref_model, None, 0, 1)
    if cls == "TinyModelWithFrontEnd":
        flat = b200asr.FlatParams(globals()[cls]())
        assert 0 < flat.tail_offset < flat.numel and flat.params[-1].shape == (6,)      # conv.weight, conv.bias at the tail
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, cls)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, losses, flat in results:
        assert losses == pytest.approx(ref_losses, rel=1e-5), rank          # mean over the GLOBAL token count
        assert torch.allclose(flat, ref_flat, rtol=1e-4, atol=1e-6), rank    # replicas stay in sync with the 1-rank run
