"""Parity on exactly what bench.py measures ("next" #1): the default 1.75 M-parameter models of
BASELINE.json configs[1] (backward-step sampling, ONE LF input shared by S samples, f16x3) and configs[2]
(cylinder-array training: 3 levels x 16 steps, rec 64, upscale 4), compared with the pinned oracle.

Stated tolerances: fields / latents 2e-4 abs, log-dets 1e-5 rel (the fp32 tolerance of the whole suite); gradients of the
benchmarked f16x3 mode within 2e-4 of the largest entry of each parameter's gradient (measured 7.5e-5 against a float64
evaluation of the oracle); the same bound for the exact-fp32 leg except on the four tensors behind one ReLU mask bit that is
inside fp32 rounding noise on these inputs (see the comment at the tolerance)."""
import types

import pytest
import torch

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda:0")


def _bench():
    import bench
    return bench


def _cyl_model():
    """The model bench.py --workload train builds (measure_train)."""
    import contextlib
    import io
    import numpy as np
    from tmglow_b200 import TMGlow
    b = _bench()
    torch.manual_seed(12345); np.random.seed(12345)
    with contextlib.redirect_stdout(io.StringIO()):
        m = TMGlow(b.TRAIN_GEOM["nic"], b.TRAIN_GEOM["noc"], [4, 4, 4], [16, 16, 16], **b.TRAIN_KW)
    b.perturb_(m, 12346)
    return m


@pytest.mark.parametrize("precision,block", [("f16x3", True), ("f16x3", False), ("fp32", False)])
def test_default_cylinder_training_gradients(precision, block):
    """configs[2] as benchmarked: default cylinder model, x[2,3,16,16] -> y[2,3,64,64], two BPTT time steps with carried
    LSTM states, training-mode BatchNorm, TMGLowLoss -> backward: the gradient of EVERY parameter against torch autograd
    through the oracles.  Exercises what the small golden model does not: C = 48 level, rec 64 (N = 256 gate conv with two
    M tiles in the weight-gradient kernel), upscale 4, the level-2 LSTM tail."""
    from oracle import tmglow_oracle as O
    from oracle import tmglow_loss_oracle as OL
    from tmglow_b200 import loss as L
    b = _bench()
    m = _cyl_model()
    sd0 = {k: v.detach().clone() for k, v in m.state_dict().items()}
    ocfg = O.OracleConfig.from_dict(m._cfg_dict)
    B, Tn, G = 2, 2, b.TRAIN_GEOM
    gen = torch.Generator().manual_seed(31)
    x = torch.randn(B, Tn, G["nic"], G["h"], G["w"], generator=gen)
    tgt = torch.randn(B, Tn, G["noc"], G["H"], G["W"], generator=gen)
    t_mean, t_rms = OL.target_statistics(tgt)
    mu, sd3 = torch.tensor([0.1, -0.2, 0.05]), torch.tensor([0.9, 1.1, 0.7])
    dx, dy, beta = 5.0 / 64, 5.0 / 64, 200.0
    h0 = O.init_lstm_states(ocfg, torch.arange(B), [G["H"], G["W"]])
    eps = [O.draw_eps(ocfg, B, G["H"], G["W"], gen) for _ in range(Tn)]
    # oracle: fp32 autograd on the CPU
    trainable = {n for n, _ in m.named_parameters()}
    sd = {k: (v.clone().requires_grad_(True) if k in trainable else v.clone()) for k, v in sd0.items()}
    h = [(a.clone(), c.clone()) for a, c in h0]
    ys, lds = [], []
    for t in range(Tn):
        y, ld, h = O.reconstruct(sd, ocfg, x[:, t], h, eps[t], training=True)
        ys.append(y); lds.append(ld)
    ref = OL.tmglow_loss(torch.stack(ys, 1), torch.stack(lds, 1), tgt, t_rms, mu, sd3, dx, dy, beta)
    ref.backward()
    # CUDA path
    dev = _dev()
    m = m.to(dev).train()
    m.precision = precision
    crit = L.TMGLowLoss(types.SimpleNamespace(beta=beta, dx=dx, dy=dy), types.SimpleNamespace(out_mu=mu, out_std=sd3)).to(dev)
    m.zero_flat_grad()
    hh = [(a.to(dev), c.to(dev)) for a, c in h0]
    ys_c, lds_c = [], []
    if block:         # what bench.py --workload train runs: the whole BPTT block in one library call
        outs = m.reconstruct_block_train(x.to(dev), hh, [[e.to(dev) for e in eps[t]] for t in range(Tn)])
        ys_c = [outs[0][:, t] for t in range(Tn)]; lds_c = [outs[1][:, t] for t in range(Tn)]
    for t in range(Tn if not block else 0):
        outs = m.reconstruct_train(x[:, t].to(dev), hh, [e.to(dev) for e in eps[t]])
        ys_c.append(outs[0]); lds_c.append(outs[1])
        hh = [(outs[2 + 2 * l], outs[3 + 2 * l]) for l in range(len(hh))]
    e_y = max((a.detach().cpu() - r.detach()).abs().max().item() for a, r in zip(ys_c, ys))
    e_ld = max(((a.detach().cpu() - r.detach()).abs() / r.detach().abs()).max().item() for a, r in zip(lds_c, lds))
    assert e_y < 2e-4 and e_ld < 1e-5, (e_y, e_ld)
    loss = crit(torch.stack(ys_c, 1), torch.stack(lds_c, 1), tgt.to(dev), t_mean.to(dev), t_rms.to(dev))
    assert abs(loss.item() - ref.item()) <= 2e-5 * abs(ref.item())
    loss.backward()
    m.scatter_flat_grad()
    params = dict(m.named_parameters())
    # Measured on B200 against a float64 evaluation of the oracle (tools/dbg_grads.py; the fp32 oracle itself is within
    # 1.2e-5 of it): f16x3 -- the mode bench.py trains in -- worst parameter 7.5e-5; exact fp32 within 6e-5 on every parameter
    # except the four ConvLSTM tensors of the level-0 LSTM step (LSTM_out_conv.weight 2.7e-3, its bias 3e-4, convLSTM.conv
    # 7e-4 / 1e-4).  That is a ReLU kink, not arithmetic: on these inputs ONE pre-activation of LSTM_out_conv (t = 0, output
    # channel 35) is 1.6e-7 in float64 and 2.9e-7 in the fp32 oracle (tools/chk_relu_kink.py) -- inside the rounding noise of
    # any fp32 evaluation -- the fp32 CUDA path lands on the other side of zero, and the gradient of exactly that channel moves
    # by that pixel's share (bias gradient: channel 35 off by 1.0e-3 absolute, the other 37 channels by < 1e-6).  The tensors
    # downstream of that one mask bit get the looser bound, everything else the 2e-4 of the suite.
    kink = ("glow.flow_blocks.0.revlayers.affine_layer16.coupling.resid_lstm.",)
    def tol_of(k):
        return 5e-3 if (precision != "f16x3" and k.startswith(kink)) else 2e-4
    checked, worst = 0, (0.0, "")
    for k in sorted(trainable):
        if sd[k].grad is None:
            continue                                  # norm2.* of the LSTM steps: declared, never used (flowLSTMBlock.py:170)
        r = sd[k].grad
        err = (params[k].grad.cpu() - r).abs().max().item()
        rel = err / max(r.abs().max().item(), 1e-3)
        worst = max(worst, (rel, k))
        assert rel <= tol_of(k), "%s: %.3e vs max %.3e" % (k, err, r.abs().max().item())
        checked += 1
    print("default cylinder model, %s: %d parameters checked, worst relative gradient error %.2e (%s)" % ((precision, checked) + worst))
    assert checked >= 530, checked            # 545 trainable tensors, 6 of them (norm2.*) unused


@pytest.mark.parametrize("S", [512])
def test_shared_input_many_samples_vs_oracle(S):
    """configs[1] as benchmarked: ONE LF snapshot shared by S stochastic samples (hoisted conditioning tables, several tiles /
    sample groups per CTA in every fused kernel), f16x3, two chained time steps; 8 scattered samples are compared with the
    oracle evaluated on exactly those samples (explicit noise, own LSTM states)."""
    from oracle import tmglow_oracle as O
    b = _bench()
    m = b.build_model()
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    ocfg = O.OracleConfig.from_dict(m._cfg_dict)
    G = b.GEOM
    gen = torch.Generator().manual_seed(5)
    x1 = torch.randn(1, G["nic"], G["h"], G["w"], generator=gen)
    x2 = torch.randn(1, G["nic"], G["h"], G["w"], generator=gen)
    pick = torch.tensor([0, 1, 77, 150, 255, 256, 300, S - 1])
    h0 = O.init_lstm_states(ocfg, torch.arange(S), [G["H"], G["W"]])
    shapes = O.latent_shapes(ocfg, S, G["H"], G["W"])
    eps = [[torch.randn(s, generator=gen) for s in shapes] for _ in range(2)]
    dev = _dev()
    m = m.to(dev).eval()
    m.precision = "f16x3"
    hh = [(a.to(dev), c.to(dev)) for a, c in h0]
    outs = []
    for t, xx in enumerate((x1, x2)):
        y, ld, hh = m.reconstruct(xx.to(dev).expand(S, -1, -1, -1), hh, [e.to(dev) for e in eps[t]])
        outs.append((y[pick.to(dev)].cpu(), ld[pick.to(dev)].cpu()))
    h = [(a[pick].clone(), c[pick].clone()) for a, c in h0]
    with torch.no_grad():
        for t, xx in enumerate((x1, x2)):
            y_o, ld_o, h = O.reconstruct(sd, ocfg, xx.expand(len(pick), -1, -1, -1).contiguous(), h, [e[pick] for e in eps[t]])
            e_y = (outs[t][0] - y_o).abs().max().item()
            e_l = ((outs[t][1] - ld_o).abs() / ld_o.abs()).max().item()
            assert e_y < 2e-4 and e_l < 1e-5, (t, e_y, e_l)
    for l, (a, c) in enumerate(hh):
        assert (a[pick.to(dev)].cpu() - h[l][0]).abs().max().item() < 2e-4
        assert (c[pick.to(dev)].cpu() - h[l][1]).abs().max().item() < 2e-4


def test_fp16_operand_overflow_is_flagged_not_silent():
    """The fp16-operand modes clamp staged activations to +-6e4 (the reference carries them in fp32): the clamp must not be
    silent.  Well-conditioned inputs leave the sticky flag clear; a latent blown up far beyond the fp16 range raises it, is
    reported through tmg_last_error (check_overflow -> FloatingPointError) and clears on request."""
    import json
    from conftest import load_golden
    from tmglow_b200 import TMGlow
    g = load_golden("caseA_states")
    cfg = json.loads(g["config"])
    m = TMGlow(cfg["in_features"], cfg["out_features"], cfg["enc_blocks"], cfg["glow_blocks"],
               cond_features=cfg["cond_features"], cglow_upscale=cfg["cglow_upscale"], growth_rate=cfg["growth_rate"],
               init_features=cfg["init_features"], rec_features=cfg["rec_features"])
    m.load_state_dict(g["state_dict"])
    dev = _dev()
    m = m.to(dev).eval()
    m.precision = "f16x3"
    h_in = [(a.to(dev), c.to(dev)) for a, c in g["h_in"]]
    eps = [e.to(dev) for e in g["rec2"]["eps"]]
    m.reconstruct(g["x"].to(dev), h_in, eps)
    assert m.activation_overflow() is False
    m.reconstruct(g["x"].to(dev), h_in, [e * 1e7 for e in eps])
    assert m.activation_overflow(clear=False) is True
    with pytest.raises(FloatingPointError):
        m.check_overflow()
    assert m.activation_overflow() is False          # cleared
