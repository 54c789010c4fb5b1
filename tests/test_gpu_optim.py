"""FlatAdam.fused_step (csrc/optim.cu: gradient norm + clip + weight decay on the trainable entries + AMSGrad, two launches,
no host sync) against the reference trainer's optimizer arithmetic in plain PyTorch: torch.nn.utils.clip_grad_norm_ followed
by torch.optim.Adam(weight_decay, amsgrad=True).step() (nn/trainFlowParallel.py:290-291, main.py:78) on the same flat tensors.
Tolerance: 1e-5 of the accumulated parameter update after 5 steps (the norm is accumulated in fp64 here, fp32 in torch)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("max_norm,wd,amsgrad", [(1.0, 1e-8, True), (None, 0.0, True), (0.05, 1e-2, False)])
def test_fused_step_matches_torch_adam(max_norm, wd, amsgrad):
    from tmglow_b200 import TMGlow, FlatAdam
    dev = torch.device("cuda:0")
    m = TMGlow(4, 3, [2, 2], [3, 3], cond_features=8, cglow_upscale=2, growth_rate=4, init_features=8, rec_features=8).to(dev)
    flat = m.flat_parameter_for_optimizer()
    mask = m.trainable_mask()
    opt = FlatAdam(m, lr=3e-3, amsgrad=amsgrad)
    # reference: the same flat tensor, masked decay, torch's own clip + Adam
    ref = torch.nn.Parameter(flat.detach().clone())
    ropt = torch.optim.Adam([ref], lr=3e-3, amsgrad=amsgrad)
    gen = torch.Generator().manual_seed(3)
    before = flat.detach().clone()
    for it in range(5):
        g = (torch.randn(flat.numel(), generator=gen) * (0.5 + it)).to(dev) * mask       # buffers never receive gradient
        out = opt.fused_step(g.clone(), max_norm=max_norm, weight_decay=wd, mask=mask)
        ref.grad = g.clone()
        norm = float(ref.grad.norm())
        if max_norm is not None:
            torch.nn.utils.clip_grad_norm_([ref], max_norm)
        ref.grad.addcmul_(ref.detach(), mask, value=wd)
        ropt.step()
        # Adam moves entries with zero gradient too (bias-corrected moments stay 0 -> no step), except through wd: masked
        with torch.no_grad():
            ref.copy_(torch.where(mask > 0, ref, before))
        assert abs(float(out[0]) - norm) <= 1e-5 * norm
        if it == 1:
            opt.param_groups[0]["lr"] = 1e-3; ropt.param_groups[0]["lr"] = 1e-3          # a scheduler changes the lr between steps
    a, r = flat.detach(), ref.detach()
    assert torch.equal(a[mask == 0], before[mask == 0])                                  # permutations, masks, BN statistics
    # parameters move by ~lr per step; the two evaluations differ by fp32 rounding of the update (bias corrections in fp32
    # here, in Python doubles in torch; gradient norm in fp64 here, fp32 in torch): 1e-5 of the accumulated update
    upd = (r - before).abs().max().item()
    err = (a - r).abs().max().item()
    assert err <= 1e-5 * upd + 5e-7, (err, upd)            # + two fp32 ulps of an O(1) parameter
    st = opt.state[flat]
    assert float(st["step"]) == 5.0 and (("max_exp_avg_sq" in st) == amsgrad)
    rs = ropt.state[ref]
    assert (st["exp_avg"] - rs["exp_avg"])[mask > 0].abs().max().item() <= 2e-6 * rs["exp_avg"].abs().max().item()


def test_flat_adam_checkpoint_round_trip(tmp_path):
    """FlatAdam is a torch.optim.Adam: the workspace conversion to the reference's per-parameter layout applies unchanged, and
    after load_state_dict the device-side step counter follows the file."""
    import types
    from tmglow_b200 import TMGlow, FlatAdam, workspace as W
    dev = torch.device("cuda:0")
    m = TMGlow(4, 3, [2, 2], [3, 3], cond_features=8, cglow_upscale=2, growth_rate=4, init_features=8, rec_features=8).to(dev)
    opt = FlatAdam(m, lr=1e-3)
    mask = m.trainable_mask()
    g = torch.randn(mask.numel(), device=dev) * mask
    for _ in range(3):
        opt.fused_step(g, max_norm=1.0, weight_decay=1e-8, mask=mask)
    args = types.SimpleNamespace(ckpt_dir=str(tmp_path), device="cpu", epoch_start=0, epochs=1, lr=1e-3)
    W.saveWorkspace(args, m, opt, file_id=1)
    _, msd, osd = W.loadWorkspace(args, str(tmp_path), file_id=1)
    assert len(osd["state"]) == len(list(m.parameters())) and float(osd["state"][0]["step"]) == 3.0
    m2 = TMGlow(4, 3, [2, 2], [3, 3], cond_features=8, cglow_upscale=2, growth_rate=4, init_features=8, rec_features=8)
    m2.load_state_dict(msd)
    m2 = m2.to(dev)
    opt2 = FlatAdam(m2, lr=1.0)
    W.load_flat_optimizer_state(m2, opt2, osd)
    opt2.fused_step(g, max_norm=1.0, weight_decay=1e-8, mask=m2.trainable_mask())
    opt.fused_step(g, max_norm=1.0, weight_decay=1e-8, mask=mask)
    assert float(opt2.state[m2.flat_parameter_for_optimizer()]["step"]) == 4.0
    assert torch.allclose(m2.flat_parameters(), m.flat_parameters(), rtol=1e-6, atol=1e-8)
