"""Generate the committed golden fixtures by RUNNING THE REAL REFERENCE.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

It imports ``nn.tmGlow.TMGlow`` from /root/reference/tmglow (unmodified), builds a few small
model configurations, perturbs the zero-initialised parameters so that couplings and priors are
non-trivial, runs the reference's own ``forward`` / ``reconstruct`` /
sub-module ``forward``/``reverse`` under ``torch.no_grad()`` and stores weights, inputs and
outputs as ``tests/golden/<name>.pt`` (plain dict of tensors, loadable with weights_only=True).
Nothing in the repo reads /root/reference at test or bench time.
"""
import copy
import json
import os
import sys
import warnings

import numpy as np
import torch

REF = "/root/reference/tmglow"
HERE = os.path.dirname(os.path.abspath(__file__))


def perturb(model, zc, gen):
    """Non-degenerate, well-conditioned weights (SURVEY.md Appendix B recipe)."""
    with torch.no_grad():
        for name, p in model.named_parameters():
            r = lambda: torch.randn(p.shape, generator=gen)
            if name.endswith("norm.weight") or name.endswith("norm2.weight"):
                p.copy_(torch.exp(0.1 * r()))
            elif name.endswith("norm.bias") or name.endswith("norm2.bias"):
                p.copy_(0.1 * r())
            elif name.endswith("conv.log_s"):
                p.add_(0.05 * r())
            elif name.endswith(".scale"):
                p.copy_(0.1 * r())
            elif "zero_conv.conv." in name or "latent_encoder.conv2d.conv." in name:
                p.copy_(zc * r())
        # non-trivial BatchNorm statistics / affine
        for name, b in model.named_buffers():
            if name.endswith("running_mean"):
                b.copy_(0.1 * torch.randn(b.shape, generator=gen))
            elif name.endswith("running_var"):
                b.copy_(1.0 + 0.2 * torch.rand(b.shape, generator=gen))
        for name, p in model.named_parameters():
            if "norm1.weight" in name:
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=gen))
            elif "norm1.bias" in name:
                p.copy_(0.1 * torch.randn(p.shape, generator=gen))


def flat_states(lst):
    return [None if s is None else (s[0].clone(), s[1].clone()) for s in lst]


def build_case(name, cfg, xshape, B, zc, seed, with_states, train_bn=False, modules=False):
    from nn.tmGlow import TMGlow
    torch.manual_seed(seed)
    np.random.seed(seed)
    model = TMGlow(cfg["in_features"], cfg["out_features"], cfg["enc_blocks"], cfg["glow_blocks"],
                   cond_features=cfg["cond_features"], cglow_upscale=cfg["cglow_upscale"],
                   growth_rate=cfg["growth_rate"], init_features=cfg["init_features"],
                   rec_features=cfg["rec_features"])
    gen = torch.Generator().manual_seed(seed + 1)
    perturb(model, zc, gen)
    model.train(train_bn)
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}

    h, w = xshape
    H, W = h * cfg["cglow_upscale"], w * cfg["cglow_upscale"]
    x = torch.randn(B, cfg["in_features"], h, w, generator=gen)
    y = torch.randn(B, cfg["out_features"], H, W, generator=gen)
    h_in = model.initLSTMStates(torch.arange(B) + 7, [H, W]) if with_states else None

    out = {"config": json.dumps(cfg), "train_bn": train_bn, "state_dict": sd0, "x": x, "y": y,
           "h_in": None if h_in is None else flat_states(h_in)}
    with torch.no_grad():
        z, logp, h_out, eps = model.forward(x, y, copy.deepcopy(h_in), return_eps=True)
        out["fwd"] = {"z": z.clone(), "logp": logp.clone(), "h_out": flat_states(h_out),
                      "eps": [e.clone() for e in eps]}
        if train_bn:   # running statistics after ONE forward call (the encoder ran once)
            out["fwd"]["bn_after"] = {k: v.clone() for k, v in model.state_dict().items()
                                      if "running_" in k or "num_batches" in k}
            model.load_state_dict(sd0)
        yr, ld, h_out2 = model.reconstruct(x, copy.deepcopy(h_in), [e.clone() for e in eps])
        out["rec"] = {"y": yr.clone(), "log_det": ld.clone(), "h_out": flat_states(h_out2)}
        if train_bn:
            model.load_state_dict(sd0)
        # reconstruct with fresh noise (what sample() computes, with the noise made explicit)
        eps2 = [torch.randn(e.shape, generator=gen) for e in eps]
        ys, lds, h_out3 = model.reconstruct(x, copy.deepcopy(h_in), [e.clone() for e in eps2])
        out["rec2"] = {"eps": eps2, "y": ys.clone(), "log_det": lds.clone(), "h_out": flat_states(h_out3)}
        if train_bn:
            model.load_state_dict(sd0)

        if modules:
            model.eval()
            mods = {}
            z_out, c_out = model.encoder.forward(x)
            mods["encoder"] = {"z_out": z_out.clone(), "c_out": [c.clone() for c in c_out]}
            # squeeze on a ragged-valued tensor (bit-exact permutation)
            sq = model.glow.flow_blocks[0].squeeze
            t = torch.randn(B, 5, 6, 10, generator=gen)
            mods["squeeze"] = {"x": t, "y": sq.forward(t).clone()}
            mods["unsqueeze"] = {"y": sq.forward(t).clone(), "x": sq.reverse(sq.forward(t)).clone()}
            # every step of block 0, both directions, on fresh inputs
            blk = model.glow.flow_blocks[0]
            C = cfg["out_features"] * 4
            hh, ww = H // 2, W // 2
            steps = []
            names = list(blk.revlayers._modules.keys())
            for i, nm in enumerate(names):
                layer = blk.revlayers._modules[nm]
                a = torch.randn(B, C, hh, ww, generator=gen)
                cond = torch.randn(B, cfg["cond_features"], hh, ww, generator=gen)
                rec = {"name": nm, "x": a, "cond": cond}
                if i == len(names) - 1:
                    st = (2 * torch.rand(B, cfg["rec_features"], hh, ww, generator=gen) - 1,
                          torch.randn(B, cfg["rec_features"], hh, ww, generator=gen))
                    o, ldd, so = layer.forward(a.clone(), cond.clone(), (st[0].clone(), st[1].clone()))
                    r, ldr, sr = layer.reverse(a.clone(), cond.clone(), (st[0].clone(), st[1].clone()))
                    o0, ld0, so0 = layer.forward(a.clone(), cond.clone(), None)
                    rec.update({"state": st, "fwd_state": (so[0].clone(), so[1].clone()),
                                "rev_state": (sr[0].clone(), sr[1].clone()),
                                "fwd_nostate": o0.clone(), "fwd_nostate_logdet": ld0.clone(),
                                "fwd_nostate_state": (so0[0].clone(), so0[1].clone())})
                else:
                    o, ldd = layer.forward(a.clone(), cond.clone())
                    r, ldr = layer.reverse(a.clone(), cond.clone())
                rec.update({"fwd": o.clone(), "fwd_logdet": ldd.clone() * torch.ones(B),
                            "rev": r.clone(), "rev_logdet": ldr.clone() * torch.ones(B)})
                steps.append(rec)
            mods["steps"] = steps
            # split, both directions
            zz = torch.randn(B, C, hh, ww, generator=gen)
            z1, lp, e = blk.split.forward(zz.clone(), return_eps=True)
            zr, lpr = blk.split.reverse(z1.clone(), e.clone())
            mods["split"] = {"z": zz, "z1": z1.clone(), "logp": lp.clone(), "eps": e.clone(),
                             "rev_z": zr.clone(), "rev_logp": lpr.clone()}
            # 1x1 weights
            conv = blk.revlayers._modules[names[0]].conv
            mods["conv1x1"] = {"W": conv.weight().clone(), "Winv": conv.inv_weight().clone()}
            out["modules"] = mods
    path = os.path.join(HERE, name + ".pt")
    torch.save(out, path)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB, logp/px = "
          f"{(logp / y[0].numel()).tolist()}, |y_rec - y|max = {(yr - y).abs().max().item():.2e}")


def default_layout():
    """Key/shape/dtype list of the DEFAULT model's state_dict (873 entries) for the drop-in test."""
    from nn.tmGlow import TMGlow
    torch.manual_seed(0)
    np.random.seed(0)
    m = TMGlow(4, 3, [4, 4, 4], [16, 16, 16], cond_features=32, cglow_upscale=2, growth_rate=4,
               init_features=16, rec_features=64)
    lay = [[k, list(v.shape), str(v.dtype)] for k, v in m.state_dict().items()]
    nparam = sum(p.numel() for p in m.parameters())
    with open(os.path.join(HERE, "default_state_dict_layout.json"), "w") as f:
        json.dump({"n_parameters": nparam, "entries": lay}, f)
    print("default layout:", len(lay), "entries,", nparam, "parameters")


if __name__ == "__main__":
    assert os.path.isdir(REF), "the reference is only available in the build container"
    sys.path.insert(0, REF)
    warnings.filterwarnings("ignore")
    A = dict(in_features=4, out_features=3, enc_blocks=[2, 2], glow_blocks=[3, 3], cond_features=8,
             cglow_upscale=2, growth_rate=4, init_features=8, rec_features=8)
    Bc = dict(in_features=3, out_features=3, enc_blocks=[1, 2, 1], glow_blocks=[2, 4, 1], cond_features=4,
              cglow_upscale=4, growth_rate=2, init_features=8, rec_features=4)
    Cc = dict(in_features=1, out_features=1, enc_blocks=[4, 4], glow_blocks=[4, 4], cond_features=8,
              cglow_upscale=1, growth_rate=4, init_features=48, rec_features=2)
    build_case("caseA_states", A, (8, 16), 2, 0.02, 11, with_states=True, modules=True)
    build_case("caseA_nostate", A, (8, 16), 2, 0.02, 12, with_states=False)
    build_case("caseA_trainbn", A, (8, 16), 3, 0.02, 13, with_states=True, train_bn=True)
    build_case("caseB_up4", Bc, (8, 8), 1, 0.02, 14, with_states=True)
    build_case("caseC_up1", Cc, (16, 16), 2, 0.02, 15, with_states=False)
    default_layout()
