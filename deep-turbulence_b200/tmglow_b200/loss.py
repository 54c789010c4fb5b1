"""``TMGLowLoss`` -- the reverse-KL training loss of TM-Glow behind the reference's own interface
(``nn/trainFlowParallel.py:104-177``): ``TMGLowLoss(args, model)(yPred, logp, target, target_mean, target_rms)``.

The whole loss -- un-normalisation, pressure-Poisson and divergence residuals (``pc/physicsConstrained.py``) with the
Sobel-type 3x3 filters (``pc/grad1Filter.py``, ``pc/grad2Filter.py``), MSE, RMS-of-fluctuation mismatch, entropy term
-- and its gradient w.r.t. ``yPred`` / ``logp`` are ONE fused CUDA kernel (``csrc/loss.cu``, ``tmg_tmglow_loss``)
plus a fixed-order finish kernel.  No CPU fallback.
"""
import torch
from torch import nn

from . import _lib


def _stream(device):
    return torch.cuda.current_stream(device).cuda_stream


class _LossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y_pred, logp, target, target_rms, out_mu, out_std, dx, dy, beta, want_terms):
        if y_pred.device.type != "cuda":
            raise RuntimeError("tmglow_b200.TMGLowLoss runs only on CUDA devices (no CPU fallback)")
        assert y_pred.dim() == 5 and y_pred.shape[2] == 3, "yPred must be [batch, time-steps, 3, nx, ny]"
        assert target.shape == y_pred.shape, "target must have the shape of yPred"
        B, T, _, H, W = y_pred.shape
        assert tuple(target_rms.shape) == (B, 3, H, W), "target_rms must be [batch, 3, nx, ny]"
        dev = y_pred.device
        lib = _lib.load()
        f = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
        y, lp, tg, tr, mu, sd = f(y_pred), f(logp), f(target), f(target_rms), f(out_mu).view(-1), f(out_std).view(-1)
        assert mu.numel() == 3 and sd.numel() == 3
        need_gy = ctx.needs_input_grad[0]
        need_gl = ctx.needs_input_grad[1]
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        terms = torch.empty(5, dtype=torch.float32, device=dev)
        g_y = torch.empty_like(y) if need_gy else None
        g_lp = torch.empty_like(lp) if need_gl else None
        ws = torch.empty(max(lib.tmg_tmglow_loss_workspace_bytes(B, T, H, W), 16), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.tmg_tmglow_loss(y.data_ptr(), lp.data_ptr(), tg.data_ptr(), tr.data_ptr(), mu.data_ptr(), sd.data_ptr(),
                                           B, T, H, W, lp.numel(), float(dx), float(dy), float(beta), loss.data_ptr(),
                                           terms.data_ptr(), g_y.data_ptr() if need_gy else None,
                                           g_lp.data_ptr() if need_gl else None, ws.data_ptr(), ws.numel(), _stream(dev)))
        ctx.g_y, ctx.g_lp = g_y, g_lp
        ctx.shapes = (y_pred.shape, logp.shape)
        ctx.mark_non_differentiable(terms)
        return loss.view(()), terms

    @staticmethod
    def backward(ctx, g_loss, _g_terms):
        g_y = None if ctx.g_y is None else (ctx.g_y * g_loss).view(ctx.shapes[0])
        g_lp = None if ctx.g_lp is None else (ctx.g_lp * g_loss).view(ctx.shapes[1])
        return g_y, g_lp, None, None, None, None, None, None, None, None


def tmglow_loss(y_pred, logp, target, target_rms, out_mu, out_std, dx, dy, beta, return_terms=False):
    """Functional form; returns the scalar loss (and ``[vPres, vDiv, vL1, vRMS, neg_entropy]`` when asked)."""
    loss, terms = _LossFn.apply(y_pred, logp, target, target_rms, out_mu, out_std, dx, dy, beta, True)
    return (loss, terms) if return_terms else loss


class TMGLowLoss(nn.Module):
    """Drop-in for the reference class of the same name (``nn/trainFlowParallel.py:104-153``).

    Args:
        args: object with ``beta``, ``dx``, ``dy`` (``args.py:35-37,61-63``)
        model: the TM-Glow model, wrapped (``model.module``, as the reference passes it) or bare; its ``out_mu`` /
            ``out_std`` un-normalise the prediction for the PDE residuals (``:119-120``)
    """

    def __init__(self, args, model, log=None):
        super().__init__()
        self.beta = args.beta
        self.dx, self.dy = args.dx, args.dy
        core = getattr(model, "module", model)
        self.register_buffer("output_std", torch.as_tensor(core.out_std, dtype=torch.float32).clone().view(1, -1, 1, 1))
        self.register_buffer("output_mu", torch.as_tensor(core.out_mu, dtype=torch.float32).clone().view(1, -1, 1, 1))
        self.last_terms = None

    def forward(self, yPred, logp, target, target_mean, target_rms):
        """``target_mean`` is accepted for interface parity; the reference computes ``targetMeanHat`` from it and never
        uses it (``:139``)."""
        loss, terms = _LossFn.apply(yPred, logp, target, target_rms, self.output_mu, self.output_std, self.dx, self.dy,
                                    self.beta, True)
        self.last_terms = terms
        return loss


def target_statistics(target0):
    """``trainFlowParallel.py:244-245``: time mean and RMS of the fluctuation over the full target series
    ``[B,Tmax,3,H,W]`` (host-side plumbing, once per mini-batch)."""
    mean = torch.mean(target0, dim=1)
    rms = torch.sqrt(torch.mean((target0 - mean.unsqueeze(1)) ** 2, dim=1))
    return mean, rms
