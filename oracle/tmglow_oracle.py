"""CPU oracle for the TM-Glow flow hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT.

A functional restatement (torch CPU fp32/fp64 operators, no nn.Module state) of the
conditional Glow stack of zabaras/deep-turbulence, driven only by a reference
``state_dict``.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this file; the
product path (``deep-turbulence_b200``) never does.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the real reference from
``/root/reference/tmglow`` in the build container, runs it on seeded inputs and commits the
inputs/weights/outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this
file against those vectors (max-abs difference 0.0 expected; 1e-6 allowed).

All ``file:line`` citations are relative to ``/root/reference/tmglow``.
Tensors are NCHW like the reference.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

LOG2PI = float(math.log(2.0 * math.pi))   # flowUtils.py:155
LOG5 = float(math.log(5.0))               # flowUtils.py:163,270
LOG4 = float(math.log(4.0))               # flowUtils.py:247


@dataclass
class OracleConfig:
    """Constructor arguments of ``TMGlow`` (nn/tmGlow.py:336-338) that shape the path."""
    in_features: int
    out_features: int
    enc_blocks: Sequence[int]
    glow_blocks: Sequence[int]
    cond_features: int = 8
    cglow_upscale: int = 1
    growth_rate: int = 4
    init_features: int = 48
    rec_features: int = 8
    bn_eps: float = 1e-5          # nn.BatchNorm2d default (denseBlock.py:49)
    bn_momentum: float = 0.1

    @staticmethod
    def from_dict(d: dict) -> "OracleConfig":
        keys = OracleConfig.__dataclass_fields__.keys()
        return OracleConfig(**{k: d[k] for k in keys if k in d})


SD = Dict[str, torch.Tensor]
State = Optional[Tuple[torch.Tensor, torch.Tensor]]


# ----------------------------------------------------------------------------- encoder
def _bn(sd: SD, pre: str, x: torch.Tensor, training: bool, cfg: OracleConfig,
        stats_out: Optional[dict]) -> torch.Tensor:
    """BatchNorm2d of ``_DenseLayer`` (denseBlock.py:49).  Train mode normalises with the
    biased batch variance and moves the running stats with the unbiased one."""
    w, b = sd[pre + ".weight"], sd[pre + ".bias"]
    if training:
        mean = x.mean(dim=(0, 2, 3))
        var = x.var(dim=(0, 2, 3), unbiased=False)
        if stats_out is not None:
            n = x.numel() // x.shape[1]
            m = cfg.bn_momentum
            stats_out[pre + ".running_mean"] = (1 - m) * sd[pre + ".running_mean"] + m * mean
            stats_out[pre + ".running_var"] = (1 - m) * sd[pre + ".running_var"] + m * var * (n / max(n - 1, 1))
            stats_out[pre + ".num_batches_tracked"] = sd[pre + ".num_batches_tracked"] + 1
    else:
        mean, var = sd[pre + ".running_mean"], sd[pre + ".running_var"]
    inv = torch.rsqrt(var + cfg.bn_eps)
    return (x - mean[None, :, None, None]) * (inv * w)[None, :, None, None] + b[None, :, None, None]


def upsample_bilinear(x: torch.Tensor, factor: int) -> torch.Tensor:
    """``UpsamplingLinear`` (modules/misc.py:34): bilinear, align_corners=True (also at factor 1)."""
    return F.interpolate(x, scale_factor=factor, mode="bilinear", align_corners=True)


def encoder_forward(sd: SD, cfg: OracleConfig, x: torch.Tensor, training: bool = False,
                    stats_out: Optional[dict] = None) -> Tuple[torch.Tensor, List[torch.Tensor]]:
    """``Encoder.forward`` (nn/tmGlow.py:104-129) incl. ``first_encoding`` (:131-158),
    ``enconding_transition`` (:161-186) and ``DenseBlock`` (denseBlock.py:15-100)."""
    p = "encoder."
    o = F.conv2d(x, sd[p + "first_encoder.In_conv.weight"], padding=1)
    o = F.conv2d(F.relu(o), sd[p + "first_encoder.In_conv3.weight"], stride=2, padding=1)
    c_out = []
    for i, nl in enumerate(cfg.enc_blocks):
        bp = f"{p}encoding_blocks.{i}."
        if i > 0:
            o = F.conv2d(F.relu(o), sd[f"{bp}encode_conv{i}.conv1.weight"], stride=2, padding=1)
        for l in range(1, nl + 1):
            lp = f"{bp}encode_dense_block{i}.denselayer{l}."
            y = F.conv2d(F.relu(_bn(sd, lp + "norm1", o, training, cfg, stats_out)),
                         sd[lp + "conv1.weight"], padding=1)
            o = torch.cat([o, y], 1)
        c0 = F.conv2d(o, sd[f"{p}cond_convs.{i}.0.weight"], padding=1)
        c_out.append(upsample_bilinear(c0, cfg.cglow_upscale))
    z_out = upsample_bilinear(F.conv2d(o, sd[p + "out_conv.0.weight"], padding=1), cfg.cglow_upscale)
    return z_out, c_out


# ----------------------------------------------------------------------------- flow pieces
def squeeze_fwd(x: torch.Tensor) -> torch.Tensor:
    """``CheckerSqueeze.forward`` (flowUtils.py:99-121): out[:,k*C+c,i,j]=in[:,c,2i+dr_k,2j+dc_k],
    (dr,dc)_k = (0,0),(1,0),(1,1),(0,1)."""
    return torch.cat([x[:, :, 0::2, 0::2], x[:, :, 1::2, 0::2], x[:, :, 1::2, 1::2], x[:, :, 0::2, 1::2]], 1)


def squeeze_rev(y: torch.Tensor) -> torch.Tensor:
    """``CheckerSqueeze.reverse`` (flowUtils.py:124-145)."""
    B, C, H, W = y.shape
    c0 = C // 4
    x = y.new_zeros(B, c0, 2 * H, 2 * W)
    x[:, :, 0::2, 0::2] = y[:, 0 * c0:1 * c0]
    x[:, :, 1::2, 0::2] = y[:, 1 * c0:2 * c0]
    x[:, :, 1::2, 1::2] = y[:, 2 * c0:3 * c0]
    x[:, :, 0::2, 1::2] = y[:, 3 * c0:4 * c0]
    return x


def conv1x1_weight(sd: SD, pre: str) -> torch.Tensor:
    """``InvertibleConv1x1LU.weight`` (glowConv.py:151-161): W = P (L*mask+I) (U*mask+diag(sign*e^log_s)+0.01 I)."""
    eye = sd[pre + "eye"]
    l = sd[pre + "l"] * sd[pre + "l_mask"] + eye
    u = sd[pre + "u"] * sd[pre + "u_mask"] + torch.diag(sd[pre + "log_s"].exp() * sd[pre + "sign_s"]) + 0.01 * eye
    return sd[pre + "p"] @ (l @ u)


def conv1x1_inv_weight(sd: SD, pre: str) -> torch.Tensor:
    """``InvertibleConv1x1LU.inv_weight`` (glowConv.py:163-174): U^-1 L^-1 P^-1."""
    eye = sd[pre + "eye"]
    l = sd[pre + "l"] * sd[pre + "l_mask"] + eye
    u = sd[pre + "u"] * sd[pre + "u_mask"] + torch.diag(sd[pre + "log_s"].exp() * sd[pre + "sign_s"]) + 0.01 * eye
    return u.inverse() @ (l.inverse() @ sd[pre + "p"].inverse())


def conv1x1_apply(x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """``F.conv2d(x, W.view(C,C,1,1))`` (glowConv.py:193-194,219-220)."""
    return F.conv2d(x, w.view(*w.shape, 1, 1))


def zero_conv(sd: SD, pre: str, x: torch.Tensor) -> torch.Tensor:
    """``Conv2dZeros.forward`` (flowUtils.py:238-247): replicate-pad 1, valid 3x3 conv + bias,
    times exp(clamp(scale,-4,ln4))."""
    y = F.conv2d(F.pad(x, (1, 1, 1, 1), mode="replicate"), sd[pre + "conv.weight"], sd[pre + "conv.bias"])
    return y * torch.exp(torch.clamp(sd[pre + "scale"], -4.0, LOG4))


def _dense2(sd: SD, pre: str, u: torch.Tensor) -> torch.Tensor:
    """``NoNormDenseBlock`` with 2 layers, growth 1 (denseBlock.py:135-138,149-150; flowAffine.py:49-53)."""
    for l in (1, 2):
        u = torch.cat([u, F.conv2d(F.relu(u), sd[f"{pre}denselayer{l}.conv1.weight"], padding=1)], 1)
    return u


def coupling_nn_plain(sd: SD, pre: str, a1: torch.Tensor, cond: torch.Tensor) -> torch.Tensor:
    """Coupling network of ``AffineCouplingLayer`` (flowAffine.py:49-55,73): dense block -> ReLU -> Conv2dZeros."""
    u = _dense2(sd, pre + "coupling_nn.dense_block.", torch.cat([a1, cond], 1))
    return zero_conv(sd, pre + "coupling_nn.zero_conv.", F.relu(u))


def coupling_nn_lstm(sd: SD, pre: str, a1: torch.Tensor, cond: torch.Tensor, state: State,
                     rec: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Coupling network of ``LSTMAffineCouplingLayer`` (flowAffine.py:147-159,177-191) with
    ``ResidLSTMBlock``/``ConvLSTMCell`` (convLSTM.py:54-85,137-152).  Gate order i,f,o,g."""
    t = torch.cat([a1, cond], 1)
    if state is None:                                   # convLSTM.py:66-68,98-102
        h = t.new_zeros(t.shape[0], rec, t.shape[2], t.shape[3])
        c = torch.zeros_like(h)
    else:
        h, c = state
    lp = pre + "resid_lstm."
    gates = F.conv2d(torch.cat([t, h], 1), sd[lp + "convLSTM.conv.weight"], sd[lp + "convLSTM.conv.bias"], padding=1)
    gi, gf, go, gg = torch.split(gates, rec, dim=1)
    c_next = torch.sigmoid(gf) * c + torch.sigmoid(gi) * torch.tanh(gg)
    h_next = torch.sigmoid(go) * torch.tanh(c_next)
    u = F.relu(F.conv2d(torch.cat([t, h_next], 1), sd[lp + "out_seq.LSTM_out_conv.weight"],
                        sd[lp + "out_seq.LSTM_out_conv.bias"], padding=1))
    u = _dense2(sd, pre + "dense_nn.dense_block.", u)
    hh = zero_conv(sd, pre + "out_conv.zero_conv.", F.relu(u))
    return hh, h_next, c_next


def _shift_logscale(h: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """flowAffine.py:76-77: shift = h[:,0::2]; log-scale a = 2*softsign(h[:,1::2])."""
    return h[:, 0::2], 2.0 * F.softsign(h[:, 1::2])


def flow_step_fwd(sd: SD, pre: str, x: torch.Tensor, cond: torch.Tensor, kind: str,
                  state: State = None, rec: int = 0):
    """One step y->z.  kind in {'unnormed','plain','lstm'}:
    flowLSTMBlock.py:116-129 / 53-69 / 180-198; ActNorm.forward actNorm.py:52-69;
    InvertibleConv1x1LU.forward glowConv.py:176-195 (train_sampling=True -> W^-1, logdet -sum(log_s)*HW);
    coupling flowAffine.py:59-83 / 161-198."""
    B, C, H, W = x.shape
    ld = x.new_zeros(())
    if kind != "unnormed":
        w, b = sd[pre + "norm.weight"], sd[pre + "norm.bias"]
        x = w * x + b
        ld = ld + w.abs().log().sum() * (H * W)
    x = conv1x1_apply(x, conv1x1_inv_weight(sd, pre + "conv."))
    ld = ld - sd[pre + "conv.log_s"].sum() * (H * W)
    x1, x2 = x.chunk(2, 1)
    s_out = None
    if kind == "lstm":
        h, hn, cn = coupling_nn_lstm(sd, pre + "coupling.", x1, cond, state, rec)
        s_out = (hn, cn)
    else:
        h = coupling_nn_plain(sd, pre + "coupling.", x1, cond)
    shift, a = _shift_logscale(h)
    scale = a.exp()
    x2 = (x2 + shift) * scale
    ld = ld + torch.abs(scale).log().reshape(B, -1).sum(1)
    return torch.cat([x1, x2], 1), ld, s_out


def flow_step_rev(sd: SD, pre: str, y: torch.Tensor, cond: torch.Tensor, kind: str,
                  state: State = None, rec: int = 0):
    """One step z->y (flowLSTMBlock.py:71-86 / 132-146 / 200-218).  Note every logdet has the
    SAME sign as in the forward direction (actNorm.py:81-83, glowConv.py:206-215, flowAffine.py:107)."""
    B, C, H, W = y.shape
    y1, y2 = y.chunk(2, 1)
    s_out = None
    if kind == "lstm":
        h, hn, cn = coupling_nn_lstm(sd, pre + "coupling.", y1, cond, state, rec)
        s_out = (hn, cn)
    else:
        h = coupling_nn_plain(sd, pre + "coupling.", y1, cond)
    shift, a = _shift_logscale(h)
    scale = a.exp()
    y2 = y2 / scale - shift
    ld = torch.abs(scale).log().reshape(B, -1).sum(1)
    y = conv1x1_apply(torch.cat([y1, y2], 1), conv1x1_weight(sd, pre + "conv."))
    ld = ld - sd[pre + "conv.log_s"].sum() * (H * W)
    if kind != "unnormed":
        w, b = sd[pre + "norm.weight"], sd[pre + "norm.bias"]
        y = (y - b) / w
        ld = ld + w.abs().log().sum() * (H * W)
    return y, ld, s_out


def _latent_prior(sd: SD, pre: str, z1: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """``LatentEncoder.forward`` (flowUtils.py:265-276) + ``GaussianDiag.__init__`` clamp (:163)."""
    h = F.hardtanh(zero_conv(sd, pre + "latent_encoder.conv2d.", z1), -2.0, LOG5)
    mean, logsd = h.chunk(2, 1)
    return mean, logsd.clamp(-10.0, LOG5)


def gaussian_logprob(x: torch.Tensor, mean: torch.Tensor, logsd: torch.Tensor) -> torch.Tensor:
    """``GaussianDiag.likelihood``/``log_prob`` (flowUtils.py:167-192)."""
    like = -0.5 * (LOG2PI + logsd * 2.0 + (x - mean) ** 2 / (logsd * 2.0).exp())
    return like.reshape(x.shape[0], -1).sum(1)


def split_fwd(sd: SD, pre: str, z: torch.Tensor):
    """``Split.forward`` (flowUtils.py:292-314)."""
    z1, z2 = z.chunk(2, 1)
    mean, logsd = _latent_prior(sd, pre, z1)
    return z1, gaussian_logprob(z2, mean, logsd), (z2 - mean) / logsd.exp()


def split_rev(sd: SD, pre: str, z1: torch.Tensor, eps: torch.Tensor):
    """``Split.reverse`` (flowUtils.py:316-335) with explicit eps (``GaussianDiag.sample`` :194-209)."""
    mean, logsd = _latent_prior(sd, pre, z1)
    z2 = mean + logsd.exp() * eps
    return torch.cat([z1, z2], 1), gaussian_logprob(z2, mean, logsd)


def _step_kind(s: int, n: int) -> str:
    """flowLSTMBlock.py:258-274: step 1 unnormed, steps 2..n-1 plain, step n LSTM."""
    if s == n:
        return "lstm"
    return "unnormed" if s == 1 else "plain"


def block_fwd(sd: SD, cfg: OracleConfig, b: int, x: torch.Tensor, cond: torch.Tensor, state: State):
    """``LSTMFLowBlock.forward`` (flowLSTMBlock.py:280-321)."""
    n = cfg.glow_blocks[b]
    bp = f"glow.flow_blocks.{b}."
    x = squeeze_fwd(x)
    ld = 0.0
    s_out = None
    for s in range(1, n + 1):
        x, d, so = flow_step_fwd(sd, f"{bp}revlayers.affine_layer{s}.", x, cond, _step_kind(s, n), state, cfg.rec_features)
        ld = ld + d
        s_out = so if so is not None else s_out
    z1, lp, eps = split_fwd(sd, bp + "split.", x)
    return z1, ld + lp, s_out, eps


def block_rev(sd: SD, cfg: OracleConfig, b: int, z1: torch.Tensor, cond: torch.Tensor, state: State,
              eps: torch.Tensor):
    """``LSTMFLowBlock.reverse`` (flowLSTMBlock.py:323-361)."""
    n = cfg.glow_blocks[b]
    bp = f"glow.flow_blocks.{b}."
    y, ld = split_rev(sd, bp + "split.", z1, eps)
    s_out = None
    for s in range(n, 0, -1):
        y, d, so = flow_step_rev(sd, f"{bp}revlayers.affine_layer{s}.", y, cond, _step_kind(s, n), state, cfg.rec_features)
        ld = ld + d
        s_out = so if so is not None else s_out
    return squeeze_rev(y), ld, s_out


# ----------------------------------------------------------------------------- model entry points
def _top_prior(z_out: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """tmGlow.py:399-401 / 433-434: chunk + the in-place clamp of GaussianDiag (flowUtils.py:163)."""
    cmean, clog = z_out.chunk(2, 1)
    return cmean, clog.clamp(-10.0, LOG5)


def forward(sd: SD, cfg: OracleConfig, x: torch.Tensor, y: torch.Tensor, h_in: Optional[list] = None,
            return_eps: bool = False, training: bool = False, stats_out: Optional[dict] = None):
    """``TMGlow.forward`` (nn/tmGlow.py:378-414) + ``LSTMCFlowDecoder.forward`` (:231-267).
    eps0 uses the clamped top log-std because of the reference's in-place aliasing."""
    z_out, c_out = encoder_forward(sd, cfg, x, training, stats_out)
    cmean, clog = _top_prior(z_out)
    z, ld, h_out, eps = y, 0.0, [], []
    for b in range(len(cfg.glow_blocks)):
        z, d, s, e = block_fwd(sd, cfg, b, z, c_out[b], None if h_in is None else h_in[b])
        ld = ld + d
        h_out.append(s)
        eps.append(e)
    eps.append((z - cmean) / clog.exp())
    logp = gaussian_logprob(z, cmean, clog) + ld
    return z, logp, h_out, (eps if return_eps else None)


def reconstruct(sd: SD, cfg: OracleConfig, x: torch.Tensor, h_in: Optional[list], eps: list,
                training: bool = False, stats_out: Optional[dict] = None):
    """``TMGlow.reconstruct`` (nn/tmGlow.py:442-467) + ``LSTMCFlowDecoder.reverse`` (:269-303).
    log_det excludes the top prior's log-prob, exactly like the reference."""
    z_out, c_out = encoder_forward(sd, cfg, x, training, stats_out)
    cmean, clog = _top_prior(z_out)
    yv = cmean + clog.exp() * eps[-1]
    nb = len(cfg.glow_blocks)
    ld, h_out = 0.0, [None] * nb
    for b in range(nb - 1, -1, -1):
        yv, d, s = block_rev(sd, cfg, b, yv, c_out[b], None if h_in is None else h_in[b], eps[b])
        ld = ld + d
        h_out[b] = s
    return yv, ld, h_out


def latent_shapes(cfg: OracleConfig, B: int, H: int, W: int) -> List[Tuple[int, ...]]:
    """Shapes of eps[0..nb-1] (split latents) and eps[nb] (top latent) for HF size HxW."""
    shapes, c = [], cfg.out_features
    for b in range(len(cfg.glow_blocks)):
        c, H, W = c * 4, H // 2, W // 2
        shapes.append((B, c // 2, H, W))
        c = c // 2
    shapes.append((B, c, H, W))
    return shapes


def draw_eps(cfg: OracleConfig, B: int, H: int, W: int, generator: Optional[torch.Generator] = None) -> list:
    """Noise in the order ``TMGlow.sample`` consumes it (top, then blocks nb-1..0:
    tmGlow.py:435, flowUtils.py:206), returned indexed like ``reconstruct``'s eps list."""
    shapes = latent_shapes(cfg, B, H, W)
    eps = [None] * len(shapes)
    for i in [len(shapes) - 1] + list(range(len(shapes) - 2, -1, -1)):
        eps[i] = torch.randn(shapes[i], generator=generator)
    return eps


def sample(sd: SD, cfg: OracleConfig, x: torch.Tensor, h_in: Optional[list] = None,
           generator: Optional[torch.Generator] = None):
    """``TMGlow.sample`` (nn/tmGlow.py:417-440)."""
    # encoder level i lives at h/2^(i+1)*upscale, flow level i at H/2^(i+1)  =>  H = h*upscale
    H, W = x.shape[2] * cfg.cglow_upscale, x.shape[3] * cfg.cglow_upscale
    return reconstruct(sd, cfg, x, h_in, draw_eps(cfg, x.shape[0], H, W, generator))


def init_lstm_states(cfg: OracleConfig, seeds: torch.Tensor, input_dim: Sequence[int]) -> list:
    """``TMGlow.initLSTMStates`` (nn/tmGlow.py:481-509): per-sample CPU generators, h~U(-1,1), c~N(0,1)."""
    out = []
    for i in range(len(cfg.glow_blocks)):
        hs, cs = [], []
        for j in range(seeds.shape[0]):
            g = torch.Generator().manual_seed(int(seeds[j].item()))
            dims = [1, cfg.rec_features, input_dim[0] // 2 ** (i + 1), input_dim[1] // 2 ** (i + 1)]
            hs.append(2 * torch.rand(dims, generator=g) - 1)
            cs.append(torch.randn(dims, generator=g))
        out.append((torch.cat(hs, 0), torch.cat(cs, 0)))
    return out
