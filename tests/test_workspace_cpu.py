"""Checkpoint ("workspace") compatibility with the reference (utils/utils.py:55-148), CPU only.
tests/golden/nsWorkspace7.zip was written by the REAL reference's saveWorkspace (tests/golden/make_golden_workspace.py):
reference model + Adam(model.parameters(), amsgrad) after two optimizer steps."""
import os
import types
import zipfile

import torch

from conftest import GOLDEN

CFG = dict(in_features=4, out_features=3, enc_blocks=[2, 2], glow_blocks=[3, 3], cond_features=8, cglow_upscale=2,
           growth_rate=4, init_features=8, rec_features=8)


def _model():
    from tmglow_b200 import TMGlow
    return TMGlow(CFG["in_features"], CFG["out_features"], CFG["enc_blocks"], CFG["glow_blocks"], cond_features=CFG["cond_features"],
                  cglow_upscale=CFG["cglow_upscale"], growth_rate=CFG["growth_rate"], init_features=CFG["init_features"],
                  rec_features=CFG["rec_features"])


def _args(d):
    return types.SimpleNamespace(ckpt_dir=str(d), device="cpu", epoch_start=0, epochs=3, lr=5e-4)


def test_reference_workspace_loads_into_flat_optimizer(tmp_path):
    """A zip written by the reference: arguments (black-list respected), model state_dict and the per-parameter Adam state
    scattered into the flat-parameter optimizer used by the CUDA training path."""
    from tmglow_b200 import workspace as W
    args = _args(tmp_path)
    out = W.loadWorkspace(args, GOLDEN, file_id=7)
    assert out is not None and W.loadWorkspace(args, GOLDEN, file_id=8) is None
    args, msd, osd = out
    assert args.beta == 200 and args.notes == "golden" and args.lr == 1e-3
    assert args.epoch_start == 0 and args.epochs == 3 and args.ckpt_dir == str(tmp_path)        # PARAM_BLACKLIST (utils.py:20)
    m = _model()
    m.load_state_dict(msd)                       # strict: identical keys / shapes
    names = [n for n, _ in m.named_parameters()]
    assert len(osd["state"]) == len(names) == len(osd["param_groups"][0]["params"])
    opt = torch.optim.Adam([m.flat_parameter_for_optimizer()], lr=1.0, amsgrad=True)
    _, wd = W.load_flat_optimizer_state(m, opt, osd)
    g = opt.param_groups[0]
    # the file's weight decay (main.py:78) is handed back for train_block; the flat optimizer itself must run without it
    assert abs(g["lr"] - 1e-3 * 0.995 ** 7) < 1e-12 and g["amsgrad"] and g["weight_decay"] == 0.0
    assert wd == 1e-8 and opt.reference_weight_decay == 1e-8
    st = opt.state[m.flat_parameter_for_optimizer()]
    assert float(st["step"]) == 2.0
    table = {name: (off, numel) for name, off, numel, shape in m._table}
    mask = m.trainable_mask()
    for k in ("exp_avg", "exp_avg_sq", "max_exp_avg_sq"):
        for i, n in enumerate(names):
            off, numel = table[n]
            assert torch.equal(st[k][off:off + numel], osd["state"][i][k].reshape(-1)), (k, n)
        assert float((st[k] * (1 - mask)).abs().max()) == 0.0          # buffers (masks, permutations, BN statistics): no moments
    # the flat parameter holds the loaded weights
    off, numel = table[names[3]]
    assert torch.equal(m.flat_parameters()[off:off + numel], msd[names[3]].reshape(-1))


def test_resumed_optimizer_leaves_buffers_untouched(tmp_path):
    """Resuming from a reference checkpoint must not move the non-trainable entries of the flat buffer (permutations, masks,
    BatchNorm running statistics): with the file's weight decay inside Adam they drifted by ~lr per step."""
    from tmglow_b200 import workspace as W
    args, msd, osd = W.loadWorkspace(_args(tmp_path), GOLDEN, file_id=7)
    m = _model()
    m.load_state_dict(msd)
    fp = m.flat_parameter_for_optimizer()
    opt = torch.optim.Adam([fp], lr=1.0, amsgrad=True)
    W.load_flat_optimizer_state(m, opt, osd)
    mask = m.trainable_mask()
    before = fp.detach().clone()
    for _ in range(20):
        fp.grad = torch.zeros_like(fp)
        opt.step()
    after = fp.detach()
    assert torch.equal(after[mask == 0], before[mask == 0])
    # and the decay written back to a file is the reference's, not the 0 the flat optimizer runs with
    assert W.flat_optimizer_state_to_reference(m, opt)["param_groups"][0]["weight_decay"] == 1e-8


def test_workspace_round_trip_in_reference_layout(tmp_path):
    """saveWorkspace from the flat-parameter optimizer writes the reference's file layout: same zip members, same state_dict
    keys, per-parameter optimizer entries equal to the ones the reference wrote."""
    from tmglow_b200 import workspace as W
    args0 = _args(tmp_path)
    args0, msd, osd = W.loadWorkspace(args0, GOLDEN, file_id=7)
    m = _model()
    m.load_state_dict(msd)
    opt = torch.optim.Adam([m.flat_parameter_for_optimizer()], lr=1.0, amsgrad=True)
    W.load_flat_optimizer_state(m, opt, osd)
    path = W.saveWorkspace(args0, m, opt, file_id=11)
    with zipfile.ZipFile(path) as z, zipfile.ZipFile(os.path.join(GOLDEN, "nsWorkspace7.zip")) as zr:
        assert sorted(z.namelist()) == ["args.json", "torchModel11.pth"]
        assert sorted(zr.namelist()) == ["args.json", "torchModel7.pth"]
    assert not os.path.exists(os.path.join(str(tmp_path), "torchModel11.pth"))        # temporaries removed, like the reference
    a2, msd2, osd2 = W.loadWorkspace(_args(tmp_path), str(tmp_path), file_id=11)
    assert list(msd2.keys()) == list(msd.keys())
    assert all(torch.equal(msd2[k], msd[k]) for k in msd)
    assert osd2["param_groups"][0]["params"] == osd["param_groups"][0]["params"]
    for i in osd["state"]:
        for k in ("exp_avg", "exp_avg_sq", "max_exp_avg_sq"):
            assert torch.equal(osd2["state"][i][k], osd["state"][i][k]) and osd2["state"][i][k].shape == osd["state"][i][k].shape
        assert float(osd2["state"][i]["step"]) == float(osd["state"][i]["step"])
    # a per-parameter optimizer (the reference's own) is stored as it is
    opt_ref = torch.optim.Adam(m.parameters(), lr=1e-3, amsgrad=True)
    opt_ref.load_state_dict(osd)
    W.saveWorkspace(args0, m, opt_ref, file_id=12)
    _, _, osd3 = W.loadWorkspace(_args(tmp_path), str(tmp_path), file_id=12)
    assert torch.equal(osd3["state"][5]["exp_avg"], osd["state"][5]["exp_avg"])
