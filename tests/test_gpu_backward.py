"""Gradient parity of the backward building blocks (exact fp32 CUDA-core kernels) against torch autograd of the same
operator on the CPU (a plain PyTorch fp32 reference of a floating-point kernel).  Tolerance: 2e-5 relative to the
largest gradient entry (fp32 summation-order differences over up to B*H*W*9 terms)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ref(x, w, b, g, relu_in, pad_replicate):
    x = x.clone().requires_grad_(True); w = w.clone().requires_grad_(True); b = b.clone().requires_grad_(True)
    u = F.relu(x) if relu_in else x
    if pad_replicate:
        y = F.conv2d(F.pad(u, (1, 1, 1, 1), mode="replicate"), w, b)
    else:
        y = F.conv2d(u, w, b, padding=1)
    y.backward(g)
    return x.grad, w.grad, b.grad


@pytest.mark.parametrize("B,Cin,Cout,H,W", [(2, 38, 1, 32, 64), (3, 40, 12, 16, 32), (2, 102, 256, 8, 16), (1, 6, 12, 5, 7),
                                            (2, 58, 48, 8, 16), (1, 3, 5, 1, 9), (2, 4, 3, 6, 1)])
@pytest.mark.parametrize("relu_in,pad_replicate", [(False, False), (True, False), (True, True)])
def test_conv3x3_backward(B, Cin, Cout, H, W, relu_in, pad_replicate):
    from tmglow_b200 import ops
    gen = torch.Generator().manual_seed(B * 1000 + Cin * 10 + Cout)
    x = torch.randn(B, Cin, H, W, generator=gen)
    w = torch.randn(Cout, Cin, 3, 3, generator=gen) * 0.1
    b = torch.randn(Cout, generator=gen)
    g = torch.randn(B, Cout, H, W, generator=gen)
    gx_r, gw_r, gb_r = _ref(x, w, b, g, relu_in, pad_replicate)
    dev = torch.device("cuda:0")
    gx, gw, gb = ops.conv3x3_backward(x.to(dev), w.to(dev), g.to(dev), relu_in, pad_replicate)
    for name, a, r in (("gx", gx, gx_r), ("gw", gw, gw_r), ("gb", gb, gb_r)):
        err = (a.cpu() - r).abs().max().item()
        assert err <= 2e-5 * max(r.abs().max().item(), 1.0), "%s: max abs err %.3e (ref max %.3e)" % (name, err, r.abs().max().item())
    # deterministic: same call twice -> bit-identical
    gx2, gw2, gb2 = ops.conv3x3_backward(x.to(dev), w.to(dev), g.to(dev), relu_in, pad_replicate)
    assert torch.equal(gx, gx2) and torch.equal(gw, gw2) and torch.equal(gb, gb2)


@pytest.mark.parametrize("B,Cin,Cout,H,W", [(2, 38, 1, 32, 64), (3, 40, 12, 16, 32), (2, 102, 256, 8, 16), (1, 6, 12, 5, 7),
                                            (2, 120, 56, 8, 16), (1, 3, 5, 1, 9), (2, 4, 3, 6, 1), (2, 58, 48, 17, 33),
                                            (4, 102, 256, 32, 32)])
@pytest.mark.parametrize("relu_in,pad_replicate", [(False, False), (True, False), (True, True)])
def test_conv3x3_wgrad_tensor_core(B, Cin, Cout, H, W, relu_in, pad_replicate):
    """Weight / bias gradient through the tcgen05 kernel (pixels as the GEMM K dimension, MN-major fp16 hi/lo operands,
    csrc/wgrad_f16.cu) against torch autograd: fp32-grade (same 2e-5 tolerance as the exact-fp32 kernel), with gradient
    magnitudes far below the fp16 range (scaled by 1e-5: the power-of-two rescaling is exercised), deterministic."""
    from tmglow_b200 import ops
    gen = torch.Generator().manual_seed(B * 1000 + Cin * 10 + Cout)
    x = torch.randn(B, Cin, H, W, generator=gen)
    w = torch.randn(Cout, Cin, 3, 3, generator=gen) * 0.1
    b = torch.randn(Cout, generator=gen)
    g = torch.randn(B, Cout, H, W, generator=gen) * 1e-5
    _, gw_r, gb_r = _ref(x, w, b, g, relu_in, pad_replicate)
    dev = torch.device("cuda:0")
    gw, gb = ops.conv3x3_wgrad_tc(x.to(dev), g.to(dev), relu_in, pad_replicate)
    for name, a, r in (("gw", gw, gw_r), ("gb", gb, gb_r)):
        err = (a.cpu() - r).abs().max().item()
        assert err <= 2e-5 * r.abs().max().item(), "%s: max abs err %.3e (ref max %.3e)" % (name, err, r.abs().max().item())
    gw2, gb2 = ops.conv3x3_wgrad_tc(x.to(dev), g.to(dev), relu_in, pad_replicate)
    assert torch.equal(gw, gw2) and torch.equal(gb, gb2)


@pytest.mark.parametrize("step", [1, 2, 3])
def test_flow_step_backward_vs_oracle_autograd(step):
    """Reverse flow step (un-normed step 1, plain steps 2-3 of block 0): gradients w.r.t. the input, the conditioning map
    and every parameter of the step against torch autograd through the pinned oracle."""
    import json
    from conftest import load_golden
    from oracle import tmglow_oracle as O
    from tmglow_b200 import TMGlow, ops
    g = load_golden("caseA_states")
    cfg = json.loads(g["config"])
    m = TMGlow(cfg["in_features"], cfg["out_features"], cfg["enc_blocks"], cfg["glow_blocks"],
               cond_features=cfg["cond_features"], cglow_upscale=cfg["cglow_upscale"], growth_rate=cfg["growth_rate"],
               init_features=cfg["init_features"], rec_features=cfg["rec_features"])
    m.load_state_dict(g["state_dict"])
    dev = torch.device("cuda:0")
    m = m.to(dev).eval()
    rec = g["modules"]["steps"][step - 1]
    x, cond = rec["x"], rec["cond"]
    gen = torch.Generator().manual_seed(step)
    g_out = torch.randn(x.shape, generator=gen)
    g_ld = torch.randn(x.shape[0], generator=gen)
    pre = "glow.flow_blocks.0.revlayers.affine_layer%d." % step
    trainable = {n for n, _ in m.named_parameters()}          # buffers (p, sign_s, masks, eye, log_s_old) get no gradient
    sd = {k: (v.clone().requires_grad_(True) if k in trainable and k.startswith(pre) else v.clone())
          for k, v in g["state_dict"].items()}
    xr = x.clone().requires_grad_(True); cr = cond.clone().requires_grad_(True)
    nsteps = cfg["glow_blocks"][0]
    kind = "unnormed" if step == 1 else ("lstm" if step == nsteps else "plain")
    state = g_state = None
    if kind == "lstm":
        R = cfg["rec_features"]
        hs = [torch.randn(x.shape[0], R, x.shape[2], x.shape[3], generator=gen).requires_grad_(True) for _ in range(2)]
        g_state = [torch.randn(x.shape[0], R, x.shape[2], x.shape[3], generator=gen) for _ in range(2)]
        y, ld, (hn, cn) = O.flow_step_rev(sd, pre, xr, cr, kind, (hs[0], hs[1]), R)
        ((y * g_out).sum() + (ld * g_ld).sum() + (hn * g_state[0]).sum() + (cn * g_state[1]).sum()).backward()
        state = [t.detach() for t in hs]
    else:
        y, ld, _ = O.flow_step_rev(sd, pre, xr, cr, kind)
        ((y * g_out).sum() + (ld * g_ld).sum()).backward()
    gx, gc, grads, gin = ops.flow_step_backward(m, 0, step, x.to(dev), cond.to(dev), g_out.to(dev), g_ld.to(dev), state, g_state)

    def close(a, r, what):
        err = (a.cpu() - r).abs().max().item()
        assert err <= 2e-5 * max(r.abs().max().item(), 1.0), "%s: max abs err %.3e (ref max %.3e)" % (what, err, r.abs().max().item())
    close(gx, xr.grad, "g_x"); close(gc, cr.grad, "g_cond")
    if kind == "lstm":
        close(gin[0], hs[0].grad, "g_h_in"); close(gin[1], hs[1].grad, "g_c_in")
    checked = 0
    for k, v in sd.items():
        if k.startswith(pre) and v.requires_grad and v.grad is not None:
            close(grads[k], v.grad, k)
            checked += 1
    assert checked >= 8, checked


@pytest.mark.parametrize("precision,rel", [("fp32", 5e-5), ("f16x3", 2e-4)])
@pytest.mark.parametrize("train_bn", [False, True])
def test_reconstruct_backward_vs_oracle_autograd(train_bn, precision, rel):
    """Whole reverse pass (TMGlow.sample with explicit noise) with BPTT over two time steps: gradients of EVERY
    parameter (encoder incl. BatchNorm in eval and train mode, flow steps, split priors) and of the initial LSTM
    states against torch autograd through the pinned oracle."""
    import json
    from conftest import load_golden
    from oracle import tmglow_oracle as O
    from tmglow_b200 import TMGlow
    g = load_golden("caseA_states")
    cfg = json.loads(g["config"])
    m = TMGlow(cfg["in_features"], cfg["out_features"], cfg["enc_blocks"], cfg["glow_blocks"],
               cond_features=cfg["cond_features"], cglow_upscale=cfg["cglow_upscale"], growth_rate=cfg["growth_rate"],
               init_features=cfg["init_features"], rec_features=cfg["rec_features"])
    m.load_state_dict(g["state_dict"])
    dev = torch.device("cuda:0")
    m = m.to(dev)
    m.train(train_bn)
    m.precision = precision      # f16x3: tensor-core forward AND tensor-core recompute inside the backward (fp32-grade)
    ocfg = O.OracleConfig.from_dict(cfg)
    x = g["x"]
    eps = g["rec2"]["eps"]
    gen = torch.Generator().manual_seed(0)
    trainable = {n for n, _ in m.named_parameters()}
    sd = {k: (v.clone().requires_grad_(True) if k in trainable else v.clone()) for k, v in g["state_dict"].items()}
    h0 = [(h.clone().requires_grad_(True), c.clone().requires_grad_(True)) for h, c in g["h_in"]]
    wy = [torch.randn(g["rec2"]["y"].shape, generator=gen) for _ in range(2)]
    wl = [torch.randn(x.shape[0], generator=gen) * 0.01 for _ in range(2)]
    # oracle: two chained time steps (same LF input and noise), loss = sum_t <y_t, wy_t> + <log_det_t, wl_t>
    h = h0
    loss = 0.0
    for t in range(2):
        y, ld, h = O.reconstruct(sd, ocfg, x, h, eps, training=train_bn)
        loss = loss + (y * wy[t]).sum() + (ld * wl[t]).sum()
    loss.backward()
    # CUDA
    m.zero_flat_grad()
    hd = [(a.detach().to(dev).requires_grad_(True), b.detach().to(dev).requires_grad_(True)) for a, b in g["h_in"]]
    hh = hd
    loss_d = 0.0
    for t in range(2):
        outs = m.reconstruct_train(x.to(dev), hh, [e.to(dev) for e in eps])
        y, ld = outs[0], outs[1]
        hh = [(outs[2 + 2 * l], outs[3 + 2 * l]) for l in range(len(hd))]
        loss_d = loss_d + (y * wy[t].to(dev)).sum() + (ld * wl[t].to(dev)).sum()
    assert abs(loss_d.item() - loss.item()) <= 1e-4 * max(1.0, abs(loss.item()))
    loss_d.backward()
    m.scatter_flat_grad()

    def close(a, r, what):
        err = (a.detach().cpu() - r).abs().max().item()
        assert err <= rel * max(r.abs().max().item(), 1.0), "%s: max abs err %.3e (ref max %.3e)" % (what, err, r.abs().max().item())
    for (a, b), (ar, br) in zip(hd, h0):
        close(a.grad, ar.grad, "g_h0"); close(b.grad, br.grad, "g_c0")
    params = dict(m.named_parameters())
    checked = 0
    for k in sorted(trainable):
        if sd[k].grad is None:          # norm2.* of the LSTM step: declared, never used (flowLSTMBlock.py:170)
            continue
        close(params[k].grad, sd[k].grad, k)
        checked += 1
    assert checked >= 60, checked


def test_train_block_reduces_loss():
    """A few optimizer steps through tmglow_b200.train.train_block (sample_train -> loss -> hand-written backward ->
    Adam on the flat parameter buffer) reduce the loss on a fixed batch; parameters stay finite and the derived
    weights follow them."""
    import json
    from conftest import load_golden
    from tmglow_b200 import TMGlow, train as T
    g = load_golden("caseA_states")
    cfg = json.loads(g["config"])
    m = TMGlow(cfg["in_features"], cfg["out_features"], cfg["enc_blocks"], cfg["glow_blocks"],
               cond_features=cfg["cond_features"], cglow_upscale=cfg["cglow_upscale"], growth_rate=cfg["growth_rate"],
               init_features=cfg["init_features"], rec_features=cfg["rec_features"])
    m.load_state_dict(g["state_dict"])
    dev = torch.device("cuda:0")
    m = m.to(dev).train()
    m.precision = "f16x3"
    gen = torch.Generator().manual_seed(1)
    B, Tn = 2, 3
    x = torch.randn(B, Tn, *g["x"].shape[1:], generator=gen).to(dev)
    tgt = 0.1 * torch.randn(B, Tn, *g["rec2"]["y"].shape[1:], generator=gen).to(dev)
    opt = torch.optim.Adam([m.flat_parameter_for_optimizer()], lr=2e-3, amsgrad=True)
    losses = []
    torch.manual_seed(3)
    for it in range(8):
        torch.manual_seed(3)                       # same noise every step: the loss is a deterministic function of the weights
        loss, norm, _ = T.train_block(m, opt, x, tgt, None, max_norm=1.0)
        losses.append(float(loss))
        assert math.isfinite(losses[-1]) and math.isfinite(norm)
    assert losses[-1] < losses[0], losses
    assert all(torch.isfinite(p).all() for p in m.parameters())
