"""CPU oracle of the TM-Glow training loss -- TEST INFRASTRUCTURE, not product code.

Functional torch-CPU restatement of ``TMGLowLoss.forward`` (reference ``nn/trainFlowParallel.py:121-177``) with the
physics residuals of ``PhysConstrainedLES`` (``pc/physicsConstrained.py:43-94``) and the smoothed finite-difference
filters ``Grad1Filter2d`` / ``Grad2Filter2d`` (``pc/grad1Filter.py:34-88``, ``pc/grad2Filter.py:33-101``), 3x3 kernels
(``trainFlowParallel.py:116``: ``grad_kernels=[3, 3]``).  Only tests/, ``__graft_entry__.smoke()`` and bench.py's CPU
legs may import this module.  **Pinned**: ``tests/golden/make_golden_loss.py`` runs the REAL reference classes and
commits inputs/outputs/gradients (``tests/golden/loss_*.pt``); ``tests/test_oracle_golden.py`` checks this restatement
against them.  Gradients come from torch autograd on this restatement (same op graph as the reference).
"""
import math

import torch
import torch.nn.functional as F

# pc/grad1Filter.py:34-36 (3x3, "weight_h"; "weight_v" is its transpose, :47)
_G1 = torch.tensor([[-1.0, 0.0, 1.0], [-2.0, 0.0, 2.0], [-1.0, 0.0, 1.0]]) / 8.0
# pc/grad2Filter.py:33-35
_G2 = torch.tensor([[1.0, -2.0, 1.0], [2.0, -4.0, 2.0], [1.0, -2.0, 1.0]]) / 4.0


def _filt(u, k, h):
    """F.conv2d(F.pad(u, 1, 'constant'), k) / h   (grad1Filter.py:60-62, grad2Filter.py:83-85). u: [N,1,H,W]."""
    return F.conv2d(F.pad(u, (1, 1, 1, 1), mode="constant"), k.to(u.dtype).view(1, 1, 3, 3)) / h


def grad1_x(u, dx):
    return _filt(u, _G1, dx)                      # grad1Filter.py:50-63


def grad1_y(u, dy):
    return _filt(u, _G1.t(), dy)                  # grad1Filter.py:65-78


def grad2_x(u, dx):
    return _filt(u, _G2, dx ** 2)                 # grad2Filter.py:73-86


def grad2_y(u, dy):
    return _filt(u, _G2.t(), dy ** 2)             # grad2Filter.py:88-101


def divergence(u_pred, dx, dy):
    """PhysConstrainedLES.calcDivergence (physicsConstrained.py:43-63): replicate the edge columns, Sobel gradients,
    scale by dx, clamp to [-1, 1].  u_pred [N,2,H,W] -> [N,1,H,W+2]."""
    up = torch.cat((u_pred[:, :, :, 0].unsqueeze(-1), u_pred, u_pred[:, :, :, -1].unsqueeze(-1)), dim=-1)
    ustar = grad1_y(up[:, 1].unsqueeze(1), dy) + grad1_x(up[:, 0].unsqueeze(1), dx)
    return torch.clamp(dx * ustar, -1, 1)


def pressure_poisson(u_pred, p_pred, dx, dy, rho=1.0):
    """PhysConstrainedLES.calcPressurePoisson (physicsConstrained.py:65-94). -> [N,1,H,W]."""
    ddp = (1.0 / rho) * (grad2_x(p_pred, dx) + grad2_y(p_pred, dy))
    u, v = u_pred[:, 0].unsqueeze(1), u_pred[:, 1].unsqueeze(1)
    rhs = grad1_x(u, dx) ** 2 + 2 * grad1_y(u, dy) * grad1_x(v, dx) + grad1_y(v, dy) ** 2
    return torch.clamp(dx * dy * (ddp + rhs), -1, 1)


def tmglow_loss(y_pred, logp, target, target_rms, out_mu, out_std, dx, dy, beta, return_terms=False):
    """TMGLowLoss.forward (trainFlowParallel.py:121-153).  y_pred, target [B,T,3,H,W]; logp [B,T] (or any shape: only
    its mean is used); target_rms [B,3,H,W]; out_mu/out_std [3].  Returns the scalar loss (and the five terms)."""
    std = out_std.view(1, -1, 1, 1)
    mu = out_mu.view(1, -1, 1, 1)
    yp = y_pred.reshape(-1, y_pred.size(-3), y_pred.size(-2), y_pred.size(-1))
    y_hat = std * yp + mu                                                         # :160, :173
    p_star = pressure_poisson(y_hat[:, :2], y_hat[:, 2:], dx, dy)                 # :162
    v_pres = torch.mean(torch.pow(p_star[:, :, 1:-1, 1:-1], 2))                   # :164
    u_star = divergence(y_hat[:, :2], dx, dy)                                     # :175
    v_div = torch.mean(torch.pow(u_star[:, :, 1:-1, 1:-1], 2))                    # :177
    v_l1 = torch.mean(torch.pow(y_pred - target, 2))                              # :142
    pred_rms = torch.sqrt(torch.mean((y_pred - torch.mean(y_pred, dim=1).unsqueeze(1)) ** 2, dim=1))   # :144
    v_rms = torch.mean(torch.pow(pred_rms - target_rms, 2))                       # :145
    n_out = y_pred.size(-3) * y_pred.size(-2) * y_pred.size(-1)
    neg_entropy = logp.mean() / math.log(2.0) / n_out                             # :148-149
    loss = beta * (v_pres + v_div + v_l1 + v_rms) + neg_entropy                   # :151
    if return_terms:
        return loss, torch.stack([v_pres, v_div, v_l1, v_rms, neg_entropy]).detach()
    return loss


def target_statistics(target0):
    """trainFlowParallel.py:244-245: time mean and RMS of the fluctuation of the FULL target series [B,Tmax,3,H,W]."""
    mean = torch.mean(target0, dim=1)
    rms = torch.sqrt(torch.mean((target0 - mean.unsqueeze(1)) ** 2, dim=1))
    return mean, rms
