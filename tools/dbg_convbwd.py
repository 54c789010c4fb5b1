import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "deep-turbulence_b200"), os.path.join(ROOT, "tests")]
import torch
import torch.nn.functional as F
from tmglow_b200 import ops
def ref(x, w, b, g, relu_in, rep):
    x = x.double().clone().requires_grad_(True); w = w.double().clone().requires_grad_(True); b = b.double().clone().requires_grad_(True)
    u = F.relu(x) if relu_in else x
    y = F.conv2d(F.pad(u, (1, 1, 1, 1), mode="replicate"), w, b) if rep else F.conv2d(u, w, b, padding=1)
    y.backward(g.double())
    return x.grad, w.grad, b.grad
dev = torch.device("cuda:0")
for (B, Cin, Cout, H, W) in [(2, 102, 38, 32, 32), (2, 102, 256, 32, 32), (2, 108, 44, 16, 16), (2, 40, 12, 32, 32), (2, 102, 38, 16, 16), (8, 102, 38, 32, 32)]:
    for relu_in, rep in [(False, False), (True, True)]:
        gen = torch.Generator().manual_seed(1)
        x = torch.randn(B, Cin, H, W, generator=gen); w = torch.randn(Cout, Cin, 3, 3, generator=gen) * 0.1
        b = torch.randn(Cout, generator=gen); g = torch.randn(B, Cout, H, W, generator=gen)
        gx_r, gw_r, gb_r = ref(x, w, b, g, relu_in, rep)
        gx, gw, gb = ops.conv3x3_backward(x.to(dev), w.to(dev), g.to(dev), relu_in, rep)
        e = lambda a, r: (a.cpu().double() - r).abs().max().item() / r.abs().max().item()
        msg = "B%d %d->%d %dx%d relu=%d rep=%d  fp32: gx %.1e gw %.1e gb %.1e" % (B, Cin, Cout, H, W, relu_in, rep, e(gx, gx_r), e(gw, gw_r), e(gb, gb_r))
        try:
            gw2, gb2 = ops.conv3x3_wgrad_tc(x.to(dev), g.to(dev), relu_in, rep)
            msg += " | tc: gw %.1e gb %.1e" % (e(gw2, gw_r), e(gb2, gb_r))
        except Exception as ex:
            msg += " | tc: " + repr(ex)[:60]
        print(msg)
