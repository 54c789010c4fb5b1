import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "deep-turbulence_b200"), os.path.join(ROOT, "tests")]
import torch
import test_gpu_bench_parity as T
from oracle import tmglow_oracle as O
from tmglow_b200 import ops
prec = sys.argv[1]; level = int(sys.argv[2]); step = int(sys.argv[3])
m = T._cyl_model()
sd0 = {k: v.detach().clone() for k, v in m.state_dict().items()}
cfg = m._cfg_dict
C = 3 * 4 * 2 ** level; Hl = 64 >> (level + 1)
B = 2
gen = torch.Generator().manual_seed(3)
x = torch.randn(B, C, Hl, Hl, generator=gen); cond = torch.randn(B, 32, Hl, Hl, generator=gen)
g_out = torch.randn(x.shape, generator=gen); g_ld = torch.randn(B, generator=gen)
pre = "glow.flow_blocks.%d.revlayers.affine_layer%d." % (level, step)
trainable = {n for n, _ in m.named_parameters()}
dt = torch.float64
sd = {k: ((v.to(dt).clone().requires_grad_(True) if k in trainable and k.startswith(pre) else v.to(dt).clone()) if v.is_floating_point() else v.clone()) for k, v in sd0.items()}
xr = x.to(dt).clone().requires_grad_(True); cr = cond.to(dt).clone().requires_grad_(True)
kind = "unnormed" if step == 1 else ("lstm" if step == 16 else "plain")
state = g_state = None
if kind == "lstm":
    R = 64
    hs = [torch.randn(B, R, Hl, Hl, generator=gen).to(dt).requires_grad_(True) for _ in range(2)]
    g_state = [torch.randn(B, R, Hl, Hl, generator=gen) for _ in range(2)]
    y, ld, (hn, cn) = O.flow_step_rev(sd, pre, xr, cr, kind, (hs[0], hs[1]), R)
    ((y * g_out.to(dt)).sum() + (ld * g_ld.to(dt)).sum() + (hn * g_state[0].to(dt)).sum() + (cn * g_state[1].to(dt)).sum()).backward()
    state = [t.detach().float() for t in hs]
else:
    y, ld, _ = O.flow_step_rev(sd, pre, xr, cr, kind)
    ((y * g_out.to(dt)).sum() + (ld * g_ld.to(dt)).sum()).backward()
dev = torch.device("cuda:0")
m = m.to(dev).eval(); m.precision = prec
gx, gc, grads, gin = ops.flow_step_backward(m, level, step, x.to(dev), cond.to(dev), g_out.to(dev), g_ld.to(dev), state, g_state)
rel = lambda a, r: (a.cpu().double() - r).abs().max().item() / max(r.abs().max().item(), 1e-6)
print(prec, "level", level, "step", step, "g_x %.1e g_cond %.1e" % (rel(gx, xr.grad), rel(gc, cr.grad)), ("g_h %.1e g_c %.1e" % (rel(gin[0], hs[0].grad), rel(gin[1], hs[1].grad))) if kind == "lstm" else "")
for k, v in sd.items():
    if k.startswith(pre) and v.requires_grad and v.grad is not None:
        print("   %.1e  %s" % (rel(grads[k], v.grad), k[len(pre):]))
