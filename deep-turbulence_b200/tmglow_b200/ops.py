"""Single operators of the flow stack through the C ABI (NCHW tensors like the reference).
They mirror the reference sub-modules one to one and are what the parity tests call:

    squeeze_forward / squeeze_reverse   CheckerSqueeze           nn/modules/flowUtils.py:99-145
    flow_step                           *CouplingBlock.forward/.reverse   nn/modules/flowLSTMBlock.py:53-218
    split_forward / split_reverse       Split                    nn/modules/flowUtils.py:292-335
"""
import torch

from . import _lib
from .nn.tmGlow import _empty_channels_last


def _stream(device):
    return torch.cuda.current_stream(device).cuda_stream


def _need_cuda(t):
    if t.device.type != "cuda":
        raise RuntimeError("tmglow_b200 operators run only on CUDA devices (no CPU fallback)")


def squeeze_forward(x):
    _need_cuda(x)
    x = x.detach().float().contiguous()
    B, C, H, W = x.shape
    assert H % 2 == 0 and W % 2 == 0
    y = torch.empty((B, 4 * C, H // 2, W // 2), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().tmg_squeeze_forward(x.data_ptr(), y.data_ptr(), B, C, H, W, _stream(x.device)))
    return y


def squeeze_reverse(y):
    _need_cuda(y)
    y = y.detach().float().contiguous()
    B, C, H, W = y.shape
    assert C >= 4 and C % 4 == 0
    x = torch.empty((B, C // 4, 2 * H, 2 * W), dtype=torch.float32, device=y.device)
    with torch.cuda.device(y.device):
        _lib.check(_lib.load().tmg_squeeze_reverse(y.data_ptr(), x.data_ptr(), B, C, H, W, _stream(y.device)))
    return x


def _op_workspace(model, lib, h, level, B, Hl, Wl, device):
    up = model._cfg.cglow_upscale
    H, W = Hl << (level + 1), Wl << (level + 1)
    assert H % up == 0 and W % up == 0
    return model._workspace(lib, h, B, H // up, W // up, device)


def flow_step(model, level, step, x, cond, state=None, reverse=False):
    """One flow step of block ``level`` (``step`` is 1-based like ``affine_layer{step}``).
    Returns ``(out, logdet[B], state_out or None)``."""
    _need_cuda(x)
    device = x.device
    lib, h = model._prepare(device)
    x = x.detach().float().contiguous()
    cond = cond.detach().float().contiguous()
    B, C, Hl, Wl = x.shape
    out = torch.empty_like(x)
    logdet = torch.empty(B, dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        st = _stream(device)
        ws = _op_workspace(model, lib, h, level, B, Hl, Wl, device)
        d = (B, model.rec_features, Hl, Wl)
        is_lstm = step == model.glow_blocks[level]
        hp = cp = None
        keep = []
        if state is not None:
            ha, ca, keep = model._states_in(lib, [state], [d], device, st, check_len=False)
            hp, cp = ha[0], ca[0]
        ho = _empty_channels_last(d, device) if is_lstm else None
        co = _empty_channels_last(d, device) if is_lstm else None
        _lib.check(lib.tmg_flow_step(h, level, step, int(reverse), B, Hl, Wl, x.data_ptr(), cond.data_ptr(), hp, cp,
                                     out.data_ptr(), logdet.data_ptr(),
                                     ho.data_ptr() if is_lstm else None, co.data_ptr() if is_lstm else None,
                                     ws.data_ptr(), ws.numel(), st))
    return out, logdet, ((ho, co) if is_lstm else None)


def split_forward(model, level, z, return_eps=True):
    _need_cuda(z)
    device = z.device
    lib, h = model._prepare(device)
    z = z.detach().float().contiguous()
    B, C, Hl, Wl = z.shape
    z1 = torch.empty((B, C // 2, Hl, Wl), dtype=torch.float32, device=device)
    eps = torch.empty_like(z1) if return_eps else None
    logp = torch.empty(B, dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        ws = _op_workspace(model, lib, h, level, B, Hl, Wl, device)
        _lib.check(lib.tmg_split_forward(h, level, B, Hl, Wl, z.data_ptr(), z1.data_ptr(), logp.data_ptr(),
                                         eps.data_ptr() if return_eps else None, ws.data_ptr(), ws.numel(),
                                         _stream(device)))
    return z1, logp, eps


def split_reverse(model, level, z1, eps):
    _need_cuda(z1)
    device = z1.device
    lib, h = model._prepare(device)
    z1 = z1.detach().float().contiguous()
    eps = eps.detach().float().contiguous()
    B, Ch, Hl, Wl = z1.shape
    z = torch.empty((B, 2 * Ch, Hl, Wl), dtype=torch.float32, device=device)
    logp = torch.empty(B, dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        ws = _op_workspace(model, lib, h, level, B, Hl, Wl, device)
        _lib.check(lib.tmg_split_reverse(h, level, B, Hl, Wl, z1.data_ptr(), eps.data_ptr(), z.data_ptr(),
                                         logp.data_ptr(), ws.data_ptr(), ws.numel(), _stream(device)))
    return z, logp


def conv3x3(x_nchw, weight, bias=None, relu_in=False, pad_replicate=False, act=0, mode="fp32"):
    """``nn.Conv2d(Cin, Cout, 3, padding=1)`` (+ input ReLU / replicate padding / activation) through the
    library's convolution kernels; ``mode`` in {"fp32" (CUDA-core FMA), "tf32x3", "tf32" (tcgen05)}.
    NCHW in/out for convenience (converted with the library's own permutation kernels)."""
    _need_cuda(x_nchw)
    device = x_nchw.device
    lib = _lib.load()
    x = x_nchw.detach().float().contiguous()
    w = weight.detach().float().contiguous()
    B, Cin, H, W = x.shape
    Cout = w.shape[0]
    assert w.shape == (Cout, Cin, 3, 3)
    b = None if bias is None else bias.detach().float().contiguous()
    xh = torch.empty((B, H, W, Cin), dtype=torch.float32, device=device)
    oh = torch.empty((B, H, W, Cout), dtype=torch.float32, device=device)
    out = torch.empty((B, Cout, H, W), dtype=torch.float32, device=device)
    ws = torch.empty(lib.tmg_conv3x3_workspace_bytes(Cin, Cout), dtype=torch.uint8, device=device)
    with torch.cuda.device(device):
        st = _stream(device)
        _lib.check(lib.tmg_nchw_to_nhwc(x.data_ptr(), xh.data_ptr(), B, Cin, H, W, st))
        _lib.check(lib.tmg_conv3x3(_lib.PRECISIONS[mode], xh.data_ptr(), B, H, W, Cin, w.data_ptr(),
                                   None if b is None else b.data_ptr(), Cout, int(relu_in), int(pad_replicate), act,
                                   oh.data_ptr(), ws.data_ptr(), ws.numel(), st))
        _lib.check(lib.tmg_nhwc_to_nchw(oh.data_ptr(), out.data_ptr(), B, Cout, H, W, st))
    return out


def conv3x3_backward(x_nchw, weight, gout_nchw, relu_in=False, pad_replicate=False):
    """Gradients of ``F.conv2d(pad(relu?(x)), weight, bias)`` w.r.t. x, weight and bias through the library's backward
    kernels (exact fp32): returns ``(gx [B,Cin,H,W], gw [Cout,Cin,3,3], gbias [Cout])``."""
    _need_cuda(x_nchw)
    device = x_nchw.device
    lib = _lib.load()
    x = x_nchw.detach().float().contiguous()
    w = weight.detach().float().contiguous()
    g = gout_nchw.detach().float().contiguous()
    B, Cin, H, W = x.shape
    Cout = w.shape[0]
    xh = torch.empty((B, H, W, Cin), dtype=torch.float32, device=device)
    gh = torch.empty((B, H, W, Cout), dtype=torch.float32, device=device)
    gxh = torch.empty((B, H, W, Cin), dtype=torch.float32, device=device)
    gx = torch.empty_like(x)
    gw = torch.empty_like(w)
    gb = torch.empty(Cout, dtype=torch.float32, device=device)
    ws = torch.empty(lib.tmg_conv3x3_backward_workspace_bytes(B, H, W, Cin, Cout), dtype=torch.uint8, device=device)
    with torch.cuda.device(device):
        st = _stream(device)
        _lib.check(lib.tmg_nchw_to_nhwc(x.data_ptr(), xh.data_ptr(), B, Cin, H, W, st))
        _lib.check(lib.tmg_nchw_to_nhwc(g.data_ptr(), gh.data_ptr(), B, Cout, H, W, st))
        _lib.check(lib.tmg_conv3x3_backward(xh.data_ptr(), B, H, W, Cin, w.data_ptr(), Cout, int(relu_in), int(pad_replicate),
                                            gh.data_ptr(), gxh.data_ptr(), gw.data_ptr(), gb.data_ptr(),
                                            ws.data_ptr(), ws.numel(), st))
        _lib.check(lib.tmg_nhwc_to_nchw(gxh.data_ptr(), gx.data_ptr(), B, Cin, H, W, st))
    return gx, gw, gb


def flow_step_backward(model, level, step, x, cond, g_out, g_logdet, state=None, g_state=None):
    """Gradients of one reverse flow step (``flow_step(..., reverse=True)``): returns ``(g_x, g_cond, grads, g_state_in)``
    where ``grads`` maps the reference parameter names of the step to their gradients.  ``state`` / ``g_state`` are the
    incoming LSTM states and the gradients w.r.t. the returned ones (LSTM step only, NCHW tensors)."""
    _need_cuda(x)
    device = x.device
    lib, h = model._prepare(device)
    x = x.detach().float().contiguous(); cond = cond.detach().float().contiguous()
    g_out = g_out.detach().float().contiguous(); g_logdet = g_logdet.detach().float().contiguous()
    B, C, Hl, Wl = x.shape
    g_x = torch.empty_like(x)
    g_cond = torch.empty_like(cond)
    flat = torch.zeros(model._n_flat, dtype=torch.float32, device=device)
    n = lib.tmg_flow_step_backward_workspace_bytes(h, level, B, Hl, Wl)
    assert n > 0, lib.tmg_last_error().decode()
    ws = torch.empty(n, dtype=torch.uint8, device=device)
    R = model.rec_features
    cl = lambda t: None if t is None else t.detach().float().to(device).contiguous(memory_format=torch.channels_last)
    hs = [cl(t) for t in (state or (None, None))]
    gs = [cl(t) for t in (g_state or (None, None))]
    gin = [_empty_channels_last((B, R, Hl, Wl), device) for _ in range(2)] if (state is not None or g_state is not None) else [None, None]
    ptr = lambda t: None if t is None else t.data_ptr()
    with torch.cuda.device(device):
        _lib.check(lib.tmg_flow_step_backward(h, level, step, B, Hl, Wl, x.data_ptr(), cond.data_ptr(), ptr(hs[0]), ptr(hs[1]),
                                              g_out.data_ptr(), g_logdet.data_ptr(), ptr(gs[0]), ptr(gs[1]),
                                              g_x.data_ptr(), g_cond.data_ptr(), ptr(gin[0]), ptr(gin[1]), flat.data_ptr(),
                                              ws.data_ptr(), ws.numel(), _stream(device)))
    pre = "glow.flow_blocks.%d.revlayers.affine_layer%d." % (level, step)
    grads = {name: flat[off:off + numel].view(shape).clone() for name, off, numel, shape in model._table if name.startswith(pre)}
    return g_x, g_cond, grads, gin


def conv3x3_wgrad_tc(x_nchw, gout_nchw, relu_in=False, pad_replicate=False):
    """Weight and bias gradient of ``F.conv2d(pad(relu?(x)), w, b)`` through the tensor-core weight-gradient kernel
    (``csrc/wgrad_f16.cu``): returns ``(gw [Cout,Cin,3,3], gbias [Cout])``."""
    _need_cuda(x_nchw)
    device = x_nchw.device
    lib = _lib.load()
    x = x_nchw.detach().float().contiguous()
    g = gout_nchw.detach().float().contiguous()
    B, Cin, H, W = x.shape
    Cout = g.shape[1]
    xh = torch.empty((B, H, W, Cin), dtype=torch.float32, device=device)
    gh = torch.empty((B, H, W, Cout), dtype=torch.float32, device=device)
    gw = torch.empty((Cout, Cin, 3, 3), dtype=torch.float32, device=device)
    gb = torch.empty(Cout, dtype=torch.float32, device=device)
    ws = torch.empty(lib.tmg_conv3x3_wgrad_tc_workspace_bytes(B, H, W, Cin, Cout), dtype=torch.uint8, device=device)
    with torch.cuda.device(device):
        st = _stream(device)
        _lib.check(lib.tmg_nchw_to_nhwc(x.data_ptr(), xh.data_ptr(), B, Cin, H, W, st))
        _lib.check(lib.tmg_nchw_to_nhwc(g.data_ptr(), gh.data_ptr(), B, Cout, H, W, st))
        _lib.check(lib.tmg_conv3x3_wgrad_tc(xh.data_ptr(), B, H, W, Cin, Cout, int(relu_in), int(pad_replicate), gh.data_ptr(),
                                            gw.data_ptr(), gb.data_ptr(), ws.data_ptr(), ws.numel(), st))
    return gw, gb
