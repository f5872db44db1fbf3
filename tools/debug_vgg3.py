import sys, torch, importlib
sys.path.insert(0, '.')
import b200asr
import torch.nn.functional as F
from oracle import asr_oracle as O
from tests.helpers import rel_err
L = b200asr._lib; lib = L.load()
st = torch.cuda.current_stream().cuda_stream
nhwc = lambda t: t.permute(0, 3, 2, 1).contiguous()
B, F_, T = 2, 23, 10
cfg = O.OracleConfig(num_layers=1, feat_extractor="vgg_cnn", freq=F_)
P = {k: v.requires_grad_(True) for k, v in O.init_params(cfg, seed=5).items() if k.startswith("conv.")}
g = torch.Generator().manual_seed(4)
x = torch.randn(B, 1, F_, T, generator=g)
h1 = F.relu(F.conv2d(x, P["conv.0.weight"], P["conv.0.bias"], padding=1)); h1.retain_grad()
h2 = F.relu(F.conv2d(h1, P["conv.2.weight"], P["conv.2.bias"], padding=1)); h2.retain_grad()
q1 = F.max_pool2d(h2, 2, 2); q1.retain_grad()
h3 = F.relu(F.conv2d(q1, P["conv.5.weight"], P["conv.5.bias"], padding=1)); h3.retain_grad()
h4 = F.relu(F.conv2d(h3, P["conv.7.weight"], P["conv.7.bias"], padding=1)); h4.retain_grad()
q2 = F.max_pool2d(h4, 2, 2)
dy = torch.randn(q2.shape, generator=g)
q2.backward(dy)
new = lambda *s: torch.full(s, float('nan'), device='cuda')
C1, C2 = 64, 128
T2, F2 = T // 2, F_ // 2
cu = lambda n: P[n].detach().cuda()
xc = x.cuda()
ws = torch.empty(9 * C2 * C2, device='cuda')
y1 = new(B, T, F_, C1); L.check(lib.b200asr_conv3x3_c1_fwd(L.ptr(xc), L.ptr(cu("conv.0.weight")), L.ptr(cu("conv.0.bias")), L.ptr(y1), B, F_, T, C1, 1, st))
y2 = new(B, T, F_, C1); L.check(lib.b200asr_conv3x3_fwd(L.ptr(y1), L.ptr(cu("conv.2.weight")), L.ptr(cu("conv.2.bias")), L.ptr(y2), L.ptr(ws), B, T, F_, C1, C1, 1, 0, st))
p1 = new(B, T2, F2, C1); L.check(lib.b200asr_maxpool2x2_fwd(L.ptr(y2), L.ptr(p1), B, T, F_, C1, st))
y3 = new(B, T2, F2, C2); L.check(lib.b200asr_conv3x3_fwd(L.ptr(p1), L.ptr(cu("conv.5.weight")), L.ptr(cu("conv.5.bias")), L.ptr(y3), L.ptr(ws), B, T2, F2, C1, C2, 1, 0, st))
y4 = new(B, T2, F2, C2); L.check(lib.b200asr_conv3x3_fwd(L.ptr(y3), L.ptr(cu("conv.7.weight")), L.ptr(cu("conv.7.bias")), L.ptr(y4), L.ptr(ws), B, T2, F2, C2, C2, 1, 0, st))
print("fwd", [("%.1e" % rel_err(a, nhwc(b.detach()))) for a, b in ((y1, h1), (y2, h2), (p1, q1), (y3, h3), (y4, h4))])
dp2 = nhwc(dy).cuda()
d4 = new(B, T2, F2, C2); L.check(lib.b200asr_maxpool2x2_bwd(L.ptr(dp2), L.ptr(y4), L.ptr(d4), B, T2, F2, C2, 1, st))
ref_d4 = nhwc((h4.grad * (h4 > 0)).detach())
print("d4 %.1e" % rel_err(d4, ref_d4), "nan", int(torch.isnan(d4).sum()))
dw7 = new(C2, C2, 3, 3); db7 = new(C2)
L.check(lib.b200asr_conv3x3_bwd_weight(L.ptr(d4), L.ptr(y3), L.ptr(dw7), L.ptr(db7), L.ptr(ws), B, T2, F2, C2, C2, 0, st))
print("dw7 %.1e db7 %.1e" % (rel_err(dw7, P["conv.7.weight"].grad), rel_err(db7, P["conv.7.bias"].grad)))
# same with the reference tensors as inputs
rd4 = ref_d4.cuda(); ry3 = nhwc(h3.detach()).cuda()
L.check(lib.b200asr_conv3x3_bwd_weight(L.ptr(rd4), L.ptr(ry3), L.ptr(dw7), L.ptr(db7), L.ptr(ws), B, T2, F2, C2, C2, 0, st))
print("dw7(ref inputs) %.1e" % rel_err(dw7, P["conv.7.weight"].grad))
tw = torch.nn.grad.conv2d_weight(h3.detach(), P["conv.7.weight"].shape, (h4.grad * (h4 > 0)).detach(), padding=1)
print("torch conv2d_weight vs autograd %.1e" % rel_err(tw, P["conv.7.weight"].grad))
