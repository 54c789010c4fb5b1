"""The time-batched BPTT block (tmg_bptt_forward / tmg_bptt_backward, TMGlow.reconstruct_block_train) against the T chained
per-time-step calls it replaces (TMGlow.reconstruct_train, itself pinned to oracle autograd in test_gpu_backward.py) and
against the oracle directly: the reference evaluates a block as `tback` successive sample() calls
(nn/trainFlowParallel.py:248-277); per-sample arithmetic is identical, only the launch structure differs.
Tolerances: outputs 2e-5 abs between the two CUDA paths (same kernels, different batch tiling), gradients 2e-4 of each
parameter's largest entry; against the oracle the suite's fp32 tolerances."""
import json

import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _model(g, dev, precision, train):
    from tmglow_b200 import TMGlow
    cfg = json.loads(g["config"])
    m = TMGlow(cfg["in_features"], cfg["out_features"], cfg["enc_blocks"], cfg["glow_blocks"],
               cond_features=cfg["cond_features"], cglow_upscale=cfg["cglow_upscale"], growth_rate=cfg["growth_rate"],
               init_features=cfg["init_features"], rec_features=cfg["rec_features"])
    m.load_state_dict(g["state_dict"])
    m = m.to(dev)
    m.train() if train else m.eval()
    m.precision = precision
    return m, cfg


@pytest.mark.parametrize("precision", ["f16x3", "fp32"])
@pytest.mark.parametrize("train_bn,with_states", [(False, True), (True, True), (True, False)])
def test_block_equals_chained_time_steps(precision, train_bn, with_states):
    g = load_golden("caseA_states")
    dev = torch.device("cuda:0")
    T = 3
    gen = torch.Generator().manual_seed(11)
    x0, eps0 = g["x"], g["rec2"]["eps"]
    B = x0.shape[0]
    xb = torch.stack([x0 * (1.0 + 0.3 * t) + 0.1 * torch.randn(x0.shape, generator=gen) for t in range(T)], 1)      # [B,T,...]
    eps = [[(e * (1.0 + 0.25 * t)).to(dev) for e in eps0] for t in range(T)]
    gy = torch.randn(B, T, *g["rec2"]["y"].shape[1:], generator=gen).to(dev)
    gld = torch.randn(B, T, generator=gen).to(dev)
    h_in = [(a.to(dev), c.to(dev)) for a, c in g["h_in"]] if with_states else None
    gst = [torch.randn(a.shape, generator=gen).to(dev) for hc in g["h_in"] for a in hc]

    def run(block):
        m, cfg = _model(g, dev, precision, train_bn)
        hs = [(a.clone().requires_grad_(True), c.clone().requires_grad_(True)) for a, c in h_in] if h_in else None
        m.zero_flat_grad()
        if block:
            outs = m.reconstruct_block_train(xb.to(dev), hs, eps)
            y, ld = outs[0], outs[1]
            hT = outs[2:]
        else:
            ys, lds, h = [], [], hs
            for t in range(T):
                o = m.reconstruct_train(xb[:, t].to(dev), h, eps[t])
                ys.append(o[0]); lds.append(o[1])
                h = [(o[2 + 2 * l], o[3 + 2 * l]) for l in range(len(cfg["glow_blocks"]))]
            y, ld = torch.stack(ys, 1), torch.stack(lds, 1)
            hT = [t_ for hc in h for t_ in hc]
        obj = (y * gy).sum() + (ld * gld).sum() + sum((a * b).sum() for a, b in zip(hT, gst))
        obj.backward()
        m.finalize_flat_grad()
        gin = [t_.grad.clone() for hc in hs for t_ in hc] if hs else []
        sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
        return y.detach(), ld.detach(), [t_.detach() for t_ in hT], m.flat_grad.clone(), gin, sd

    y1, ld1, h1, g1, gi1, sd1 = run(False)
    y2, ld2, h2, g2, gi2, sd2 = run(True)
    assert (y1 - y2).abs().max().item() < 2e-5
    assert ((ld1 - ld2).abs() / ld1.abs()).max().item() < 1e-5
    for a, b in zip(h1, h2):
        assert (a - b).abs().max().item() < 2e-5
    m, _ = _model(g, dev, precision, train_bn)
    for name, off, numel, shape in m._table:
        a, b = g1[off:off + numel], g2[off:off + numel]
        if float(a.abs().max()) == 0.0 and float(b.abs().max()) == 0.0:
            continue
        err = (a - b).abs().max().item()
        assert err <= 2e-4 * max(a.abs().max().item(), 1e-3), "%s: %.3e vs max %.3e" % (name, err, a.abs().max().item())
    for a, b in zip(gi1, gi2):
        assert (a - b).abs().max().item() <= 2e-4 * max(a.abs().max().item(), 1e-3)
    if train_bn:        # BatchNorm running statistics: updated once per time step, in order
        for k in sd1:
            if k.endswith(("running_mean", "running_var", "num_batches_tracked")):
                assert torch.allclose(sd1[k].float(), sd2[k].float(), rtol=1e-5, atol=1e-6), k


def test_block_vs_oracle_autograd():
    """The block call against torch autograd through the oracle (three chained time steps, every parameter)."""
    from oracle import tmglow_oracle as O
    g = load_golden("caseA_states")
    dev = torch.device("cuda:0")
    T = 3
    m, cfg = _model(g, dev, "f16x3", False)
    ocfg = O.OracleConfig.from_dict(cfg)
    gen = torch.Generator().manual_seed(12)
    x0, eps0 = g["x"], g["rec2"]["eps"]
    B = x0.shape[0]
    xb = torch.stack([x0 * (1.0 + 0.3 * t) for t in range(T)], 1)
    eps = [[e * (1.0 + 0.25 * t) for e in eps0] for t in range(T)]
    gy = torch.randn(B, T, *g["rec2"]["y"].shape[1:], generator=gen)
    gld = torch.randn(B, T, generator=gen)
    trainable = {n for n, _ in m.named_parameters()}
    sd = {k: (v.clone().requires_grad_(True) if k in trainable else v.clone()) for k, v in g["state_dict"].items()}
    h = [(a.clone(), c.clone()) for a, c in g["h_in"]]
    ys, lds = [], []
    for t in range(T):
        y, ld, h = O.reconstruct(sd, ocfg, xb[:, t], h, eps[t])
        ys.append(y); lds.append(ld)
    ((torch.stack(ys, 1) * gy).sum() + (torch.stack(lds, 1) * gld).sum()).backward()
    m.zero_flat_grad()
    outs = m.reconstruct_block_train(xb.to(dev), [(a.to(dev), c.to(dev)) for a, c in g["h_in"]], [[e.to(dev) for e in et] for et in eps])
    assert (outs[0].cpu() - torch.stack(ys, 1).detach()).abs().max().item() < 2e-4
    ((outs[0] * gy.to(dev)).sum() + (outs[1] * gld.to(dev)).sum()).backward()
    m.scatter_flat_grad()
    params = dict(m.named_parameters())
    checked = 0
    for k in sorted(trainable):
        if sd[k].grad is None:
            continue
        r = sd[k].grad
        err = (params[k].grad.cpu() - r).abs().max().item()
        assert err <= 2e-4 * max(r.abs().max().item(), 1e-3), "%s: %.3e vs max %.3e" % (k, err, r.abs().max().item())
        checked += 1
    assert checked >= 60
