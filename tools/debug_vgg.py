import sys, torch
sys.path.insert(0, '.')
import b200asr, importlib
from oracle import asr_oracle as O
from tests.helpers import rel_err
ops = importlib.import_module(b200asr.__name__ + ".ops")
names = ["conv.0.weight", "conv.0.bias", "conv.2.weight", "conv.2.bias", "conv.5.weight", "conv.5.bias", "conv.7.weight", "conv.7.bias"]
for (B, F_, T) in [(3, 23, 10), (2, 23, 10), (3, 23, 12), (3, 24, 10), (1, 23, 10), (3, 20, 10), (3, 23, 8), (2, 41, 24)]:
    cfg = O.OracleConfig(num_layers=1, feat_extractor="vgg_cnn", freq=F_)
    P = {k: v.requires_grad_(True) for k, v in O.init_params(cfg, seed=5).items() if k.startswith("conv.")}
    g = torch.Generator().manual_seed(4)
    x = torch.randn(B, 1, F_, T, generator=g)
    y = O.vgg_frontend(x, P)
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy)
    cs = [P[n].detach().cuda().requires_grad_(True) for n in names]
    yc = ops.VggFrontendFn.apply(x.cuda(), *cs)
    yc.backward(dy.permute(0, 3, 2, 1).contiguous().cuda())
    print((B, F_, T), "fwd %.1e" % rel_err(yc.permute(0, 3, 2, 1), y), " ".join("%s=%.1e" % (n.replace("conv.", "").replace("weight", "w").replace("bias", "b"), rel_err(c.grad, P[n].grad)) for c, n in zip(cs, names)))
