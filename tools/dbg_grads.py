"""Developer check: per-parameter gradient error of the CUDA training path on the default cylinder model (bench config[2])
against a float64 evaluation of the oracle."""
import os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "deep-turbulence_b200"), os.path.join(ROOT, "tests")]
import torch
import test_gpu_bench_parity as T
from oracle import tmglow_oracle as O
from oracle import tmglow_loss_oracle as OL
from tmglow_b200 import loss as L

prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
Tn = int(sys.argv[2]) if len(sys.argv) > 2 else 2
train_bn = (sys.argv[3] != "eval") if len(sys.argv) > 3 else True
b = T._bench()
m = T._cyl_model()
sd0 = {k: v.detach().clone() for k, v in m.state_dict().items()}
ocfg = O.OracleConfig.from_dict(m._cfg_dict)
B, G = 2, b.TRAIN_GEOM
gen = torch.Generator().manual_seed(31)
x = torch.randn(B, Tn, G["nic"], G["h"], G["w"], generator=gen)
tgt = torch.randn(B, Tn, G["noc"], G["H"], G["W"], generator=gen)
t_mean, t_rms = OL.target_statistics(tgt)
mu, sd3 = torch.tensor([0.1, -0.2, 0.05]), torch.tensor([0.9, 1.1, 0.7])
h0 = O.init_lstm_states(ocfg, torch.arange(B), [G["H"], G["W"]])
eps = [O.draw_eps(ocfg, B, G["H"], G["W"], gen) for _ in range(Tn)]
trainable = {n for n, _ in m.named_parameters()}
dt = torch.float64
sd = {k: (v.to(dt).clone().requires_grad_(True) if k in trainable else (v.to(dt) if v.is_floating_point() else v.clone())) for k, v in sd0.items()}
h = [(a.to(dt), c.to(dt)) for a, c in h0]
ys, lds = [], []
for t in range(Tn):
    y, ld, h = O.reconstruct(sd, ocfg, x[:, t].to(dt), h, [e.to(dt) for e in eps[t]], training=train_bn)
    ys.append(y); lds.append(ld)
ref = OL.tmglow_loss(torch.stack(ys, 1), torch.stack(lds, 1), tgt.to(dt), t_rms.to(dt), mu.to(dt), sd3.to(dt), 5 / 64, 5 / 64, 200.)
ref.backward()
dev = torch.device("cuda:0")
m = m.to(dev)
m.train() if train_bn else m.eval()
m.precision = prec
crit = L.TMGLowLoss(types.SimpleNamespace(beta=200., dx=5 / 64, dy=5 / 64), types.SimpleNamespace(out_mu=mu, out_std=sd3)).to(dev)
m.zero_flat_grad()
hh = [(a.to(dev), c.to(dev)) for a, c in h0]
ys_c, lds_c = [], []
for t in range(Tn):
    outs = m.reconstruct_train(x[:, t].to(dev), hh, [e.to(dev) for e in eps[t]])
    ys_c.append(outs[0]); lds_c.append(outs[1])
    hh = [(outs[2 + 2 * l], outs[3 + 2 * l]) for l in range(len(hh))]
loss = crit(torch.stack(ys_c, 1), torch.stack(lds_c, 1), tgt.to(dev), t_mean.to(dev), t_rms.to(dev))
print("loss", loss.item(), float(ref.detach()), "y err", max((a.detach().cpu().double() - r.detach()).abs().max().item() for a, r in zip(ys_c, ys)))
loss.backward()
m.scatter_flat_grad()
params = dict(m.named_parameters())
errs = []
for k in sorted(trainable):
    if sd[k].grad is None:
        continue
    r = sd[k].grad
    e = (params[k].grad.cpu().double() - r).abs().max().item() / max(r.abs().max().item(), 1e-3)
    errs.append((e, k, r.abs().max().item()))
errs.sort(reverse=True)
print("prec", prec, "T", Tn, "bn_train", train_bn)
for e, k, mx in errs[:25]:
    print("%.2e  (max %.2e)  %s" % (e, mx, k))
import collections
byk = collections.defaultdict(float)
for e, k, mx in errs:
    kind = k.split(".")[-2] + "." + k.split(".")[-1] if "revlayers" in k else k.split(".")[0]
    lvl = k.split(".")[2] if k.startswith("glow.flow_blocks") else "enc"
    byk[(lvl, kind)] = max(byk[(lvl, kind)], e)
for kk, e in sorted(byk.items(), key=lambda x: -x[1])[:20]:
    print("%s %.2e" % (kk, e))
# slice analysis of the ConvLSTM tensors: error by input-channel group and by output block
for e, k, mx in errs[:8]:
    if not k.endswith("conv.weight") and not k.endswith("LSTM_out_conv.weight"):
        continue
    r = sd[k].grad
    d = (params[k].grad.cpu().double() - r).abs()
    O_, I_ = r.shape[0], r.shape[1]
    C = ocfg.out_features * 4 * (2 ** int(k.split(".")[2])) if k.startswith("glow.flow_blocks") else 0
    print(k, tuple(r.shape), "max|g| %.3e" % r.abs().max().item())
    cuts = [0, C // 2, C // 2 + 32, I_] if I_ > C // 2 + 32 else [0, I_]
    for a0, a1 in zip(cuts[:-1], cuts[1:]):
        print("   in[%3d:%3d]  err %.3e  max|g| %.3e" % (a0, a1, d[:, a0:a1].max().item(), r[:, a0:a1].abs().max().item()))
    nb = 4 if O_ % 4 == 0 and O_ >= 64 else 1
    for q in range(nb):
        s0, s1 = q * O_ // nb, (q + 1) * O_ // nb
        print("   out[%3d:%3d] err %.3e  max|g| %.3e" % (s0, s1, d[s0:s1].max().item(), r[s0:s1].abs().max().item()))
    for tap in range(9):
        print("   tap %d err %.3e" % (tap, d[:, :, tap // 3, tap % 3].max().item()), end="")
    print()
for k in [kk for _, kk, _ in errs[:4] if kk.endswith("bias")]:
    r = sd[k].grad; g_ = params[k].grad.cpu().double()
    print(k, "err per channel:", ["%.1e" % v for v in (g_ - r).abs().tolist()[:48]])
    print("   ref:", ["%.2f" % v for v in r.tolist()[:48]])
