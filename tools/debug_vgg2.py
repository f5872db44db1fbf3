import sys, torch, importlib
sys.path.insert(0, '.')
import b200asr
import torch.nn.functional as F
from oracle import asr_oracle as O
from tests.helpers import rel_err
L = b200asr._lib; lib = L.load()
st = torch.cuda.current_stream().cuda_stream
nhwc = lambda t: t.permute(0, 3, 2, 1).contiguous()       # (B,C,F,T) -> (B,T,F,C)
for (B, F_, T) in [(2, 11, 5), (2, 11, 6), (2, 12, 5), (1, 11, 5)]:
    g = torch.Generator().manual_seed(1)
    C = 128
    x = torch.randn(B, C, F_, T, generator=g).relu().requires_grad_(True)
    w = (torch.randn(C, C, 3, 3, generator=g) * 0.05).requires_grad_(True)
    y = F.relu(F.conv2d(x, w, None, padding=1)); y.retain_grad()
    p = F.max_pool2d(y, 2, 2)
    dp = torch.randn(p.shape, generator=g)
    p.backward(dp)
    yc, xc, dpc = nhwc(y.detach()).cuda(), nhwc(x.detach()).cuda(), nhwc(dp).cuda()
    d = torch.full((B, T, F_, C), float('nan'), device='cuda')
    L.check(lib.b200asr_maxpool2x2_bwd(L.ptr(dpc), L.ptr(yc), L.ptr(d), B, T, F_, C, 1, st))
    dy_ref = nhwc((y.grad * (y > 0)).detach())
    ws = torch.empty(9 * C * C, device='cuda')
    dw = torch.empty(C, C, 3, 3, device='cuda'); db = torch.empty(C, device='cuda')
    L.check(lib.b200asr_conv3x3_bwd_weight(L.ptr(d), L.ptr(xc), L.ptr(dw), L.ptr(db), L.ptr(ws), B, T, F_, C, C, 0, st))
    dref = dy_ref.cuda()
    dw2 = torch.empty(C, C, 3, 3, device='cuda')
    L.check(lib.b200asr_conv3x3_bwd_weight(L.ptr(dref), L.ptr(xc), L.ptr(dw2), L.ptr(db), L.ptr(ws), B, T, F_, C, C, 0, st))
    dx = torch.empty(B, T, F_, C, device='cuda')
    L.check(lib.b200asr_conv3x3_bwd_data(L.ptr(dref), L.ptr(w.detach().cuda()), None, L.ptr(dx), L.ptr(ws), B, T, F_, C, C, 0, st))
    print((B, F_, T), "poolbwd %.1e nan=%d" % (rel_err(d, dy_ref), int(torch.isnan(d).sum())), "wgrad(own d) %.1e wgrad(ref d) %.1e dgrad %.1e" % (rel_err(dw, w.grad), rel_err(dw2, w.grad), rel_err(dx, nhwc(x.grad))))
