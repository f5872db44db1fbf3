"""-m gpu tests of the tcgen05 kernels against the exact-fp32 CUDA-core kernels and the CPU oracle.

precision 1 (TF32, one pass) is checked at TF32 tolerance; precision 3 (3xTF32 split) must meet fp32-grade error."""
import ctypes

import pytest
import torch

from tests.helpers import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    import b200asr
    b200asr._lib.load(check_device=True)
    return b200asr._lib


def _stream():
    return torch.cuda.current_stream().cuda_stream


SHAPES = [(128, 128, 32), (128, 128, 128), (256, 384, 512), (300, 512, 512), (77, 52, 164), (6400 // 8, 512, 5120 // 4),
          (130, 4364, 128), (513, 128, 2048), (1, 8, 4), (200, 64, 96)]


@pytest.mark.parametrize("prec,tol", [(1, 2e-3), (3, 2e-5)])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_linear_forward_tensor_core(L, M, N, K, prec, tol):
    lib = L.load()
    g = torch.Generator().manual_seed(M + N + K)
    x = torch.randn(M, K, generator=g).cuda()
    w = torch.randn(N, K, generator=g).cuda()
    b = torch.randn(N, generator=g).cuda()
    ref = (x.double() @ w.double().t() + b.double()).relu()
    y = torch.full((M, N), float("nan"), device="cuda")
    L.check(lib.b200asr_linear_fwd(L.ptr(x), L.ptr(w), L.ptr(b), L.ptr(y), M, N, K, 1, prec, _stream()), "linear_fwd")
    torch.cuda.synchronize()
    assert rel_err(y, ref) < tol


@pytest.mark.parametrize("prec,tol", [(1, 2e-3), (3, 2e-5)])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_linear_backward_tensor_core(L, M, N, K, prec, tol):
    """bwd_data exercises an MN-major B operand, bwd_weight MN-major A and B (plus split-K atomics)."""
    lib = L.load()
    g = torch.Generator().manual_seed(M * 3 + N + K)
    x = torch.randn(M, K, generator=g).cuda()
    w = torch.randn(N, K, generator=g).cuda()
    dy = torch.randn(M, N, generator=g).cuda()
    h = torch.randn(M, K, generator=g).cuda()
    dx = torch.full((M, K), float("nan"), device="cuda")
    L.check(lib.b200asr_linear_bwd_data(L.ptr(dy), L.ptr(w), L.ptr(h), L.ptr(dx), M, N, K, 0, prec, _stream()), "bwd_data")
    ref_dx = (dy.double() @ w.double()) * (h > 0)
    assert rel_err(dx, ref_dx) < tol
    dw = torch.full((N, K), float("nan"), device="cuda")
    db = torch.full((N,), float("nan"), device="cuda")
    L.check(lib.b200asr_linear_bwd_weight(L.ptr(dy), L.ptr(x), L.ptr(dw), L.ptr(db), M, N, K, 0, prec, _stream()), "bwd_weight")
    assert rel_err(dw, dy.double().t() @ x.double()) < tol
    assert rel_err(db, dy.double().sum(0)) < 1e-5
    # accumulate path
    base = torch.randn(N, K, generator=g).cuda()
    dw2 = base.clone()
    L.check(lib.b200asr_linear_bwd_weight(L.ptr(dy), L.ptr(x), L.ptr(dw2), None, M, N, K, 1, prec, _stream()), "bwd_weight acc")
    assert rel_err(dw2, base.double() + dy.double().t() @ x.double()) < tol


def test_unaligned_shapes_are_rejected_not_rerouted(L):
    lib = L.load()
    x = torch.randn(8, 161).cuda(); w = torch.randn(16, 161).cuda(); y = torch.empty(8, 16).cuda()
    rc = lib.b200asr_linear_fwd(L.ptr(x), L.ptr(w), None, L.ptr(y), 8, 16, 161, 0, 3, _stream())
    assert rc == -1 and "multiples of 4" in L.last_error()
