"""Generate the golden fixtures of the training loss by RUNNING THE REAL REFERENCE.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_loss.py

Imports ``TMGLowLoss`` (reference ``nn/trainFlowParallel.py:104-177``, which pulls ``PhysConstrainedLES`` and the
Sobel filters of ``pc/``) unmodified.  ``nn.trainFlowParallel`` also imports ``utils.viz`` -> matplotlib (absent in
this image, plotting only): an empty stand-in module is registered for the import, nothing of it is called.
Stores inputs, the loss, its five terms (recomputed with the reference's own ``calcVPres``/``calcVDiv`` and the same
expressions) and the autograd gradients w.r.t. ``yPred`` and ``logp`` as ``tests/golden/loss_<name>.pt``.
"""
import math
import os
import sys
import types

import torch

REF = "/root/reference/tmglow"
HERE = os.path.dirname(os.path.abspath(__file__))


def _import_reference():
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.gridspec", "matplotlib.colors", "matplotlib.ticker"):
        if name not in sys.modules:
            try:
                __import__(name)
            except ImportError:
                sys.modules[name] = types.ModuleType(name)
    mpl = sys.modules["matplotlib"]
    if not hasattr(mpl, "use"):                   # the stand-in: viz.py calls mpl.use('agg') / mpl.rcParams at import
        class _Anything(dict):
            def __call__(self, *a, **k):
                return self

            def __getattr__(self, k):
                return self
        mpl.__path__ = []
        mpl.use = lambda *a, **k: None
        mpl.rcParams = _Anything()
        mpl.rc = lambda *a, **k: None
        mpl.pyplot = sys.modules["matplotlib.pyplot"]
        for sub in ("pyplot", "gridspec", "colors", "ticker"):
            sys.modules["matplotlib." + sub].__getattr__ = lambda k: _Anything()
    sys.path.insert(0, REF)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from nn.trainFlowParallel import TMGLowLoss
    return TMGLowLoss


class _Args:
    def __init__(self, beta, dx, dy):
        self.beta, self.dx, self.dy = beta, dx, dy


class _Model:
    """TMGLowLoss reads model.module.out_std / out_mu (trainFlowParallel.py:119-120)."""
    def __init__(self, mu, std):
        self.module = types.SimpleNamespace(out_mu=mu, out_std=std)


def build_case(TMGLowLoss, name, B, T, H, W, dx, dy, beta, seed, amp=1.0, smooth=False):
    g = torch.Generator().manual_seed(seed)
    mu = 0.3 * torch.randn(3, generator=g)
    std = 0.5 + torch.rand(3, generator=g)
    y = amp * torch.randn(B, T, 3, H, W, generator=g)
    if smooth:      # small, smooth fields: most residuals stay inside the clamp (both clamp branches are covered)
        y = torch.nn.functional.avg_pool2d(y.view(-1, 3, H, W), 5, 1, 2).view(B, T, 3, H, W) * 0.2
    target = torch.randn(B, T, 3, H, W, generator=g)
    target_mean = target.mean(1)
    target_rms = torch.sqrt(torch.mean((target - target_mean.unsqueeze(1)) ** 2, dim=1))
    logp = 50.0 * torch.randn(B, T, generator=g)
    loss_mod = TMGLowLoss(_Args(beta, dx, dy), _Model(mu, std))
    yv = y.clone().requires_grad_(True)
    lv = logp.clone().requires_grad_(True)
    loss = loss_mod(yv, lv, target, target_mean, target_rms)
    loss.backward()
    with torch.no_grad():
        flat = y.view(-1, 3, H, W)
        v_pres = loss_mod.calcVPres(flat)
        v_div = loss_mod.calcVDiv(flat)
        v_l1 = torch.mean(torch.pow(y - target, 2))
        pred_rms = torch.sqrt(torch.mean((y - torch.mean(y, dim=1).unsqueeze(1)) ** 2, dim=1))
        v_rms = torch.mean(torch.pow(pred_rms - target_rms, 2))
        neg_entropy = logp.mean() / math.log(2.0) / (3 * H * W)
        y_hat = loss_mod.output_std * flat + loss_mod.output_mu
        p_star = loss_mod.phys.calcPressurePoisson(y_hat[:, :2], y_hat[:, 2:])
        u_star = loss_mod.phys.calcDivergence(y_hat[:, :2])
    out = {"y_pred": y, "logp": logp, "target": target, "target_rms": target_rms, "out_mu": mu, "out_std": std,
           "dx": torch.tensor(dx, dtype=torch.float64), "dy": torch.tensor(dy, dtype=torch.float64),
           "beta": torch.tensor(beta, dtype=torch.float64),
           "loss": loss.detach(), "terms": torch.stack([v_pres, v_div, v_l1, v_rms, neg_entropy]),
           "g_y": yv.grad.clone(), "g_logp": lv.grad.clone(), "p_star": p_star, "u_star": u_star,
           "clamped_frac": torch.tensor([(p_star.abs() >= 1).float().mean().item(), (u_star.abs() >= 1).float().mean().item()])}
    path = os.path.join(HERE, "loss_%s.pt" % name)
    torch.save(out, path)
    print(name, "loss=%.6f" % float(loss), "terms", [round(float(t), 6) for t in out["terms"]],
          "clamped", out["clamped_frac"].tolist(), os.path.getsize(path), "bytes")


def main():
    TMGLowLoss = _import_reference()
    torch.manual_seed(0)
    # cylinder-array constants (args.py:61-63), small grid; residuals mostly clamped (what random fields give)
    build_case(TMGLowLoss, "cyl_rand", B=2, T=4, H=16, W=24, dx=5.0 / 64, dy=5.0 / 64, beta=200.0, seed=11)
    # smooth low-amplitude fields: residuals mostly inside the clamp -> the stencil adjoints carry the gradient
    build_case(TMGLowLoss, "cyl_smooth", B=2, T=5, H=20, W=18, dx=5.0 / 64, dy=5.0 / 64, beta=200.0, seed=12, smooth=True)
    # backward-step constants (args.py:35-37), anisotropic spacing on purpose (dx != dy exercises every divisor)
    build_case(TMGLowLoss, "step_aniso", B=1, T=3, H=12, W=40, dx=2.0 / 64, dy=3.0 / 64, beta=200.0, seed=13, smooth=True)


if __name__ == "__main__":
    main()
