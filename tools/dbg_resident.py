"""Developer check of the level-resident flow kernel: default bench model, S samples of one LF input, f16x3, against the oracle."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "deep-turbulence_b200")]
import torch
import bench
from oracle import tmglow_oracle as O

S = int(sys.argv[1]) if len(sys.argv) > 1 else 8
shared = (sys.argv[2] != "distinct") if len(sys.argv) > 2 else True
m = bench.build_model()
sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
cfg = O.OracleConfig.from_dict(m._cfg_dict)
G = bench.GEOM
gen = torch.Generator().manual_seed(5)
x = torch.randn(1 if shared else S, G["nic"], G["h"], G["w"], generator=gen)
h0 = O.init_lstm_states(cfg, torch.arange(S), [G["H"], G["W"]])
eps = [torch.randn(s, generator=gen) for s in O.latent_shapes(cfg, S, G["H"], G["W"])]
dev = torch.device("cuda:0")
m = m.to(dev).eval()
m.precision = os.environ.get("PREC", "f16x3")
xx = x.to(dev).expand(S, -1, -1, -1) if shared else x.to(dev)
y, ld, hh = m.reconstruct(xx, [(a.to(dev), c.to(dev)) for a, c in h0], [e.to(dev) for e in eps])
torch.cuda.synchronize()
n = min(S, 8)
pick = torch.linspace(0, S - 1, n).long()
with torch.no_grad():
    xo = (x.expand(S, -1, -1, -1) if shared else x)[pick].contiguous()
    y_o, ld_o, h_o = O.reconstruct(sd, cfg, xo, [(a[pick], c[pick]) for a, c in h0], [e[pick] for e in eps])
print("S=%d shared=%s  |y-y_o|=%.3e  rel ld=%.3e" % (S, shared, (y[pick.to(dev)].cpu() - y_o).abs().max().item(),
      ((ld[pick.to(dev)].cpu() - ld_o).abs() / ld_o.abs()).max().item()))
