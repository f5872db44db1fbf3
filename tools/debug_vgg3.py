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
_keep = {n: P[n].detach().cuda() for n in P}
cu = lambda n: _keep[n]
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
d3 = new(B, T2, F2, C2); L.check(lib.b200asr_conv3x3_bwd_data(L.ptr(d4), L.ptr(cu("conv.7.weight")), L.ptr(y3), L.ptr(d3), L.ptr(ws), B, T2, F2, C2, C2, 0, st))
print("d3 %.1e" % rel_err(d3, nhwc((h3.grad * (h3 > 0)).detach())))
dp1 = new(B, T2, F2, C1); L.check(lib.b200asr_conv3x3_bwd_data(L.ptr(d3), L.ptr(cu("conv.5.weight")), None, L.ptr(dp1), L.ptr(ws), B, T2, F2, C1, C2, 0, st))
print("dp1 %.1e" % rel_err(dp1, nhwc(q1.grad)))
d2 = new(B, T, F_, C1); L.check(lib.b200asr_maxpool2x2_bwd(L.ptr(dp1), L.ptr(y2), L.ptr(d2), B, T, F_, C1, 1, st))
print("d2 %.1e" % rel_err(d2, nhwc((h2.grad * (h2 > 0)).detach())))
# same with the reference tensors as inputs
rd4 = ref_d4.cuda(); ry3 = nhwc(h3.detach()).cuda()
L.check(lib.b200asr_conv3x3_bwd_weight(L.ptr(rd4), L.ptr(ry3), L.ptr(dw7), L.ptr(db7), L.ptr(ws), B, T2, F2, C2, C2, 0, st))
print("dw7(ref inputs) %.1e" % rel_err(dw7, P["conv.7.weight"].grad))
tw = torch.nn.grad.conv2d_weight(h3.detach(), P["conv.7.weight"].shape, (h4.grad * (h4 > 0)).detach(), padding=1)
print("torch conv2d_weight vs autograd %.1e" % rel_err(tw, P["conv.7.weight"].grad))
diff = (d4.cpu() - ref_d4).abs()
idx = (diff > 1e-4).nonzero()
print("mismatch count", idx.shape[0], "of", diff.numel())
y4c = y4.cpu(); h4n = nhwc(h4.detach())
for k in range(min(6, idx.shape[0])):
    b_, t_, f_, c_ = idx[k].tolist()
    t2, f2 = t_ // 2, f_ // 2
    print((b_, t_, f_, c_), "ours d4", d4[b_, t_, f_, c_].item(), "ref", ref_d4[b_, t_, f_, c_].item())
    if t2 * 2 + 1 < T2 and f2 * 2 + 1 < F2:
        print("   window ours", [y4c[b_, 2 * t2 + a, 2 * f2 + bb, c_].item() for a in (0, 1) for bb in (0, 1)])
        print("   window ref ", [h4n[b_, 2 * t2 + a, 2 * f2 + bb, c_].item() for a in (0, 1) for bb in (0, 1)])
        print("   dp2", dp2[b_, t2, f2, c_].item())
print("per-batch err", [rel_err(d4[i], ref_d4[i]) for i in range(B)])
