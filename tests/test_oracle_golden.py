"""Pins oracle/tmglow_oracle.py to the vectors produced by the REAL reference
(tests/golden/make_golden.py).  CPU only."""
import json

import pytest
import torch

from oracle import tmglow_oracle as O
from conftest import load_golden

CASES = ["caseA_states", "caseA_nostate", "caseA_trainbn", "caseB_up4", "caseC_up1"]
TOL = 1e-6   # observed: 0.0 (same ATen operators in the same order)


def _close(a, b, tol=TOL):
    assert a.shape == b.shape
    assert (a - b).abs().max().item() <= tol * max(1.0, b.abs().max().item())


def _states_close(a, b):
    assert len(a) == len(b)
    for s, t in zip(a, b):
        _close(s[0], t[0]); _close(s[1], t[1])


@pytest.mark.parametrize("name", CASES)
def test_forward_matches_reference(name):
    g = load_golden(name)
    cfg = O.OracleConfig.from_dict(json.loads(g["config"]))
    stats = {}
    z, logp, h_out, eps = O.forward(g["state_dict"], cfg, g["x"], g["y"], g["h_in"], True,
                                    training=g["train_bn"], stats_out=stats)
    _close(z, g["fwd"]["z"]); _close(logp, g["fwd"]["logp"])
    _states_close(h_out, g["fwd"]["h_out"])
    for a, b in zip(eps, g["fwd"]["eps"]):
        _close(a, b)
    if g["train_bn"]:
        for k, v in g["fwd"]["bn_after"].items():
            _close(stats[k].float(), v.float())


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("which", ["rec", "rec2"])
def test_reconstruct_matches_reference(name, which):
    g = load_golden(name)
    cfg = O.OracleConfig.from_dict(json.loads(g["config"]))
    eps = g["fwd"]["eps"] if which == "rec" else g["rec2"]["eps"]
    y, ld, h_out = O.reconstruct(g["state_dict"], cfg, g["x"], g["h_in"], eps, training=g["train_bn"])
    _close(y, g[which]["y"]); _close(ld, g[which]["log_det"])
    _states_close(h_out, g[which]["h_out"])


def test_modules_match_reference():
    g = load_golden("caseA_states")
    cfg = O.OracleConfig.from_dict(json.loads(g["config"]))
    sd, m = g["state_dict"], g["modules"]
    z_out, c_out = O.encoder_forward(sd, cfg, g["x"])
    _close(z_out, m["encoder"]["z_out"])
    for a, b in zip(c_out, m["encoder"]["c_out"]):
        _close(a, b)
    assert torch.equal(O.squeeze_fwd(m["squeeze"]["x"]), m["squeeze"]["y"])          # bit-exact permutation
    assert torch.equal(O.squeeze_rev(m["unsqueeze"]["y"]), m["unsqueeze"]["x"])
    n = cfg.glow_blocks[0]
    for s, rec in enumerate(m["steps"], start=1):
        pre = f"glow.flow_blocks.0.revlayers.{rec['name']}."
        kind = O._step_kind(s, n)
        st = rec.get("state")
        o, ld, so = O.flow_step_fwd(sd, pre, rec["x"], rec["cond"], kind, st, cfg.rec_features)
        _close(o, rec["fwd"]); _close(ld * torch.ones(o.shape[0]), rec["fwd_logdet"])
        r, ldr, sr = O.flow_step_rev(sd, pre, rec["x"], rec["cond"], kind, st, cfg.rec_features)
        _close(r, rec["rev"]); _close(ldr * torch.ones(o.shape[0]), rec["rev_logdet"])
        if kind == "lstm":
            _states_close([so], [rec["fwd_state"]]); _states_close([sr], [rec["rev_state"]])
            o0, ld0, so0 = O.flow_step_fwd(sd, pre, rec["x"], rec["cond"], kind, None, cfg.rec_features)
            _close(o0, rec["fwd_nostate"]); _states_close([so0], [rec["fwd_nostate_state"]])
    sp = m["split"]
    z1, lp, e = O.split_fwd(sd, "glow.flow_blocks.0.split.", sp["z"])
    _close(z1, sp["z1"]); _close(lp, sp["logp"]); _close(e, sp["eps"])
    zr, lpr = O.split_rev(sd, "glow.flow_blocks.0.split.", sp["z1"], sp["eps"])
    _close(zr, sp["rev_z"]); _close(lpr, sp["rev_logp"])
    pre = "glow.flow_blocks.0.revlayers.affine_layer1.conv."
    _close(O.conv1x1_weight(sd, pre), m["conv1x1"]["W"]); _close(O.conv1x1_inv_weight(sd, pre), m["conv1x1"]["Winv"])


def test_invertibility_property():
    """The reference's only known-answer property (nn/tmGlow.py:511-530): reconstruct(forward(y)) == y."""
    g = load_golden("caseA_states")
    cfg = O.OracleConfig.from_dict(json.loads(g["config"]))
    z, logp, h_out, eps = O.forward(g["state_dict"], cfg, g["x"], g["y"], g["h_in"], True)
    y, ld, _ = O.reconstruct(g["state_dict"], cfg, g["x"], g["h_in"], eps)
    assert (y - g["y"]).abs().max().item() < 1e-4
    # forward.logp == reconstruct.log_det + logpdf_top(z)   
    z_out, _ = O.encoder_forward(g["state_dict"], cfg, g["x"])
    cm, cl = O._top_prior(z_out)
    assert torch.allclose(logp, ld + O.gaussian_logprob(z, cm, cl), rtol=1e-5, atol=1e-3)


def test_init_lstm_states_seeded():
    cfg = O.OracleConfig(4, 3, [2, 2], [3, 3], rec_features=8)
    a = O.init_lstm_states(cfg, torch.tensor([3, 4]), [16, 32])
    b = O.init_lstm_states(cfg, torch.tensor([3, 4]), [16, 32])
    assert a[0][0].shape == (2, 8, 8, 16) and a[1][1].shape == (2, 8, 4, 8)
    assert torch.equal(a[0][0], b[0][0]) and a[0][0].abs().max() <= 1


def test_init_lstm_states_matches_reference():
    g = load_golden("caseA_states")
    cfg = O.OracleConfig.from_dict(json.loads(g["config"]))
    B = g["x"].shape[0]
    mine = O.init_lstm_states(cfg, torch.arange(B) + 7, list(g["y"].shape[2:]))
    for (h, c), (hr, cr) in zip(mine, g["h_in"]):
        assert torch.equal(h, hr) and torch.equal(c, cr)


# ---------------------------------------------------------------- training loss (tests/golden/make_golden_loss.py)
LOSS_CASES = ["loss_cyl_rand", "loss_cyl_smooth", "loss_step_aniso"]


@pytest.mark.parametrize("name", LOSS_CASES)
def test_loss_oracle_matches_reference(name):
    """oracle/tmglow_loss_oracle.py reproduces TMGLowLoss (value, terms, residual fields, autograd gradients)."""
    from oracle import tmglow_loss_oracle as L
    g = load_golden(name)
    dx, dy, beta = float(g["dx"]), float(g["dy"]), float(g["beta"])
    y = g["y_pred"].clone().requires_grad_(True)
    lp = g["logp"].clone().requires_grad_(True)
    loss, terms = L.tmglow_loss(y, lp, g["target"], g["target_rms"], g["out_mu"], g["out_std"], dx, dy, beta, return_terms=True)
    loss.backward()
    _close(loss.detach(), g["loss"]); _close(terms, g["terms"])
    _close(y.grad, g["g_y"]); _close(lp.grad, g["g_logp"])
    flat = g["y_pred"].view(-1, *g["y_pred"].shape[2:])
    y_hat = g["out_std"].view(1, 3, 1, 1) * flat + g["out_mu"].view(1, 3, 1, 1)
    _close(L.pressure_poisson(y_hat[:, :2], y_hat[:, 2:], dx, dy), g["p_star"])
    _close(L.divergence(y_hat[:, :2], dx, dy), g["u_star"])
    # the cases cover both clamp branches
    if name == "loss_cyl_rand":
        assert 0.05 < float(g["clamped_frac"][0]) < 0.95


def test_target_statistics_matches_reference_expression():
    from oracle import tmglow_loss_oracle as L
    t = torch.randn(2, 6, 3, 4, 5, generator=torch.Generator().manual_seed(3))
    mean, rms = L.target_statistics(t)
    assert torch.equal(mean, torch.mean(t, axis=1))
    assert torch.equal(rms, torch.sqrt(torch.mean((t - mean.unsqueeze(1)) ** 2, dim=1)))


def test_oracle_equals_the_reference_copy_on_the_default_model():
    """When the unmodified reference package is available (oracle/_ref, made by oracle/make_ref.py; /root/reference in the
    build container) the oracle is checked against it directly on the default 1.75 M-parameter cylinder model -- the
    configuration bench.py trains: forward, reconstruct and the LSTM states, max-abs difference 0.0."""
    import numpy as np
    import pytest
    import torch
    from oracle import ref_loader
    from oracle import tmglow_oracle as O
    if ref_loader.ref_root() is None:
        pytest.skip("no copy of the reference package on this machine")
    import bench
    ns = ref_loader.load()
    torch.manual_seed(3); np.random.seed(3)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        ref = ns.TMGlow(bench.TRAIN_GEOM["nic"], bench.TRAIN_GEOM["noc"], [4, 4, 4], [16, 16, 16], **bench.TRAIN_KW)
    bench.perturb_(ref, 11)
    ref.eval()
    cfg = O.OracleConfig(in_features=3, out_features=3, enc_blocks=[4, 4, 4], glow_blocks=[16, 16, 16], **bench.TRAIN_KW)
    sd = {k: v.detach().clone() for k, v in ref.state_dict().items()}
    gen = torch.Generator().manual_seed(8)
    B, G = 2, bench.TRAIN_GEOM
    x = torch.randn(B, G["nic"], G["h"], G["w"], generator=gen)
    y = torch.randn(B, G["noc"], G["H"], G["W"], generator=gen)
    h = ref.initLSTMStates(torch.arange(B), [G["H"], G["W"]])
    with torch.no_grad():
        z_r, lp_r, h_r, eps_r = ref(x, y, h, return_eps=True)
        z_o, lp_o, h_o, eps_o = O.forward(sd, cfg, x, y, h, return_eps=True)
        assert torch.equal(z_r, z_o) and torch.equal(lp_r, lp_o)
        for e_r, e_o in zip(eps_r, eps_o):
            assert torch.equal(e_r, e_o)
        y_r, ld_r, h2_r = ref.reconstruct(x, h, eps_r)
        y_o, ld_o, h2_o = O.reconstruct(sd, cfg, x, h, eps_o)
        assert torch.equal(y_r, y_o) and torch.equal(ld_r, ld_o)
        for (a, c), (ao, co) in zip(h2_r, h2_o):
            assert torch.equal(a, ao) and torch.equal(c, co)
