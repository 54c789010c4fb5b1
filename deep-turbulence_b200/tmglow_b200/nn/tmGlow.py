"""Drop-in ``TMGlow`` for the B200 hot path.

Mirrors the public surface of the reference model (``tmglow/nn/tmGlow.py:305-509``): same
constructor signature, ``forward`` / ``sample`` / ``reconstruct`` / ``initLSTMStates`` /
``_num_parameters``, same attributes (``glow_blocks``, ``rec_features``, ``encoder``, ``glow``,
``in_mu`` ...), and a ``state_dict()`` with exactly the reference's keys and shapes, so reference
checkpoints load unchanged.  The compute is NOT PyTorch: every call goes through the C ABI of
``libtmglow_b200.so`` (hand-written sm_100a CUDA); the sub-modules only hold parameters.  There
is no CPU fallback -- tensors must live on a CUDA device.

``forward`` / ``sample`` / ``reconstruct`` return plain tensors (no autograd graph); training runs through
``sample_train`` / ``reconstruct_train``, whose backward is hand-written CUDA for EVERY parameter (flow and encoder,
BatchNorm batch statistics included) and accumulates into the flat gradient buffer ``flat_grad``.
"""
import math
import re
from typing import List, Optional, Sequence, Tuple

import numpy as np
import scipy.linalg
import torch
import torch.nn as nn

from operator import attrgetter

from .. import _lib

_VERSION = attrgetter("_version")

_BUFFER_SUFFIXES = ("running_mean", "running_var", "conv.p", "conv.sign_s", "conv.l_mask", "conv.u_mask",
                    "conv.eye", "conv.log_s_old")
_TOP_BUFFERS = ("in_mu", "in_std", "out_mu", "out_std")
_LU_LEAF = re.compile(r"revlayers\.affine_layer\d+\.conv\.(l|u|log_s|p|sign_s|l_mask|u_mask|eye|log_s_old)$")


def _empty_channels_last(d, device):
    """[B,C,H,W]-shaped fp32 tensor whose memory is [B,H,W,C] (what the library reads/writes)."""
    return torch.empty_strided(d, (d[1] * d[2] * d[3], 1, d[3] * d[1], d[1]), dtype=torch.float32, device=device)


class _Params(nn.Module):
    """Parameter container; its children/leaves carry the reference's names."""

    def forward(self, *a, **k):   # pragma: no cover
        raise RuntimeError("parameter container of the B200 TMGlow: call the TMGlow methods instead")


class _EncoderHandle(_Params):
    """``model.encoder``: holds the encoder weights; ``forward(x)`` runs the CUDA encoder
    (reference ``Encoder.forward``, nn/tmGlow.py:104-129) and returns ``(z_out, c_out)``."""

    def forward(self, x):
        return self._owner[0].encode(x)


def _lu_init(C, rng=np.random):
    """Random rotation + LU factors, as the reference initialises ``InvertibleConv1x1LU``
    (glowConv.py:123-147): numpy QR of a Gaussian matrix, scipy LU, sign/log of diag(U)."""
    q = np.linalg.qr(rng.randn(C, C))[0].astype(np.float32)
    p, l, u = scipy.linalg.lu(q)
    s = np.diag(u)
    return {
        "l": l.astype(np.float32), "u": np.triu(u, k=1).astype(np.float32),
        "log_s": np.log(np.abs(s)).astype(np.float32), "p": p.astype(np.float32),
        "sign_s": np.sign(s).astype(np.float32),
        "l_mask": np.tril(np.ones((C, C), np.float32), -1), "u_mask": np.triu(np.ones((C, C), np.float32), 1),
        "eye": np.eye(C, dtype=np.float32), "log_s_old": (np.log(np.abs(s)) + 1.0).astype(np.float32),
    }


class TMGlow(nn.Module):
    """Transient multi-fidelity Glow (drop-in for ``nn.tmGlow.TMGlow``, nn/tmGlow.py:305-509)."""

    def __init__(self, in_features, out_features, enc_blocks, glow_blocks,
                 cond_features=8, cglow_upscale=1, growth_rate=4, init_features=48, rec_features=8, bn_size=8,
                 drop_rate=0, bottleneck=False):
        super().__init__()
        assert len(enc_blocks) == len(glow_blocks), 'List of conditions need to be same length as flow blocks.'
        if bottleneck or drop_rate:
            raise NotImplementedError("bottleneck/drop_rate are dead options in the reference (main.py:72)")
        self.glow_blocks = list(glow_blocks)
        self.rec_features = rec_features
        self._cfg_dict = dict(in_features=in_features, out_features=out_features, enc_blocks=list(enc_blocks),
                              glow_blocks=list(glow_blocks), cond_features=cond_features,
                              cglow_upscale=cglow_upscale, growth_rate=growth_rate, init_features=init_features,
                              rec_features=rec_features)
        lib = _lib.load()
        cfg = _lib.TmgConfig()
        cfg.in_features, cfg.out_features, cfg.n_levels = in_features, out_features, len(glow_blocks)
        for i, (e, g) in enumerate(zip(enc_blocks, glow_blocks)):
            cfg.enc_blocks[i], cfg.glow_blocks[i] = e, g
        cfg.cond_features, cfg.cglow_upscale, cfg.growth_rate = cond_features, cglow_upscale, growth_rate
        cfg.init_features, cfg.rec_features = init_features, rec_features
        self._cfg = cfg
        self._handles = {}            # device index -> tmg_model*
        h = self._handle("host")
        # ---- parameter tree with the reference's names, in the library's table order
        self._table = []              # (name, offset, numel, shape)
        dims = (_lib.C.c_int64 * 4)()
        for i in range(lib.tmg_model_param_entries(h)):
            nd = lib.tmg_model_param_shape(h, i, dims)
            self._table.append((lib.tmg_model_param_name(h, i).decode(), lib.tmg_model_param_offset(h, i),
                                lib.tmg_model_param_numel(h, i), tuple(int(dims[k]) for k in range(nd))))
        self._n_flat = lib.tmg_model_param_total(h)
        self.encoder = _EncoderHandle()
        self.glow = _Params()
        object.__setattr__(self.encoder, "_owner", (self,))   # tuple: keep it out of the module tree
        self._build_tree()
        self._fix_conv_bias()
        self._flat = None
        self._ws = {}
        self.always_refresh_weights = False
        self._refreshed_for = None
        self._precision = "fp32"
        self.register_load_state_dict_post_hook(TMGlow._after_load_state_dict)
        print('Total number of parameters: {}'.format(self._num_parameters()))

    @staticmethod
    def _after_load_state_dict(module, incompatible_keys):
        module.__dict__["_layout_ok_for"] = None      # load_state_dict(assign=True) replaces the leaf tensors
        module.__dict__["_leaf_tuple"] = None
        module.__dict__["_leaf_cache"] = None

    @property
    def precision(self):
        """Arithmetic of the heavy 3x3 convolutions: ``"fp32"`` (CUDA-core FMA, exact fp32),
        ``"tf32x3"`` (tcgen05 tensor cores with the 3xTF32 split: fp32-grade accuracy), ``"tf32"``
        (tcgen05, single-pass TF32 -- the default arithmetic of stock PyTorch/cuDNN convolutions on GPU),
        ``"f16x3"`` (tcgen05, fp16 hi+lo operand split with power-of-two weight scaling: fp32-grade accuracy,
        persistent fused flow-step kernel) or ``"f16"`` (single-pass fp16 operands, fp32 accumulation)."""
        return self._precision

    @precision.setter
    def precision(self, mode):
        if mode not in _lib.PRECISIONS:
            raise ValueError("precision must be one of %s" % sorted(_lib.PRECISIONS))
        self._precision = mode

    # ------------------------------------------------------------------ construction helpers
    def _handle(self, key):
        if key not in self._handles:
            out = _lib.C.c_void_p()
            _lib.check(_lib.load().tmg_model_create(_lib.C.byref(self._cfg), _lib.C.byref(out)))
            self._handles[key] = out
        return self._handles[key]

    def __del__(self):
        try:
            lib = _lib.load()
            for h in self._handles.values():
                lib.tmg_model_destroy(h)
        except Exception:
            pass

    def _build_tree(self):
        lu_cache = {}
        self._leaves = []             # (module, attr, is_param) in table order
        for name, off, numel, shape in self._table:
            parts = name.split(".")
            mod = self
            for p in parts[:-1]:
                if p not in mod._modules:
                    mod.add_module(p, _Params())
                mod = mod._modules[p]
            leaf = parts[-1]
            t, is_buf = self._init_leaf(name, shape, lu_cache)
            if is_buf:
                mod.register_buffer(leaf, t)
            else:
                mod.register_parameter(leaf, nn.Parameter(t))
            self._leaves.append((mod, leaf, not is_buf))
            if leaf == "running_var":     # BatchNorm2d's integer counter follows running_var
                mod.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))

    @staticmethod
    def _init_leaf(name, shape, lu_cache):
        """Same initial distributions as the reference modules (it relies on PyTorch defaults)."""
        leaf = name.rsplit(".", 1)[-1]
        if name in _TOP_BUFFERS:
            return torch.zeros(shape), True                                   # tmGlow.py:371-374
        if _LU_LEAF.search(name):
            key = name.rsplit(".", 1)[0]
            if key not in lu_cache:
                lu_cache[key] = _lu_init(shape[0])
            return torch.from_numpy(lu_cache[key][leaf].copy()), leaf not in ("l", "u", "log_s")
        if leaf == "running_mean":
            return torch.zeros(shape), True
        if leaf == "running_var":
            return torch.ones(shape), True
        if "norm1." in name or name.endswith(("norm.weight", "norm.bias", "norm2.weight", "norm2.bias")):
            return (torch.ones(shape) if leaf == "weight" else torch.zeros(shape)), False   # actNorm.py:31-32
        if "zero_conv" in name or "latent_encoder.conv2d" in name:
            return torch.zeros(shape), False                                  # flowUtils.py:231-233
        if leaf == "weight":                                                  # nn.Conv2d default init
            w = torch.empty(shape)
            nn.init.kaiming_uniform_(w, a=math.sqrt(5))
            return w, False
        if leaf == "bias":                                                    # nn.Conv2d default bias init
            # fan_in of the sibling weight: O = shape[0]; the conv's input channels are not known
            # from the bias alone, so the bound is filled in by _fix_conv_bias below.
            return torch.zeros(shape), False
        raise RuntimeError("unhandled parameter " + name)

    def _fix_conv_bias(self):
        for mod in self.modules():
            w = mod._parameters.get("weight") if hasattr(mod, "_parameters") else None
            b = mod._parameters.get("bias") if hasattr(mod, "_parameters") else None
            if w is not None and b is not None and w.dim() == 4 and w.abs().sum() > 0:
                bound = 1.0 / math.sqrt(w.shape[1] * 9)
                with torch.no_grad():
                    b.uniform_(-bound, bound)

    # ------------------------------------------------------------------ flat parameter buffer
    def _leaf_modules(self):
        """The container module of every table entry, resolved through THIS object's module tree: a replica made by
        ``nn.DataParallel.replicate`` (utils/parallel.py:166-169) shares ``__dict__`` entries with the original but owns
        re-wired children holding the per-device copies, so nothing is cached across objects."""
        cache = self.__dict__.get("_leaf_cache")
        if cache is None or cache[0] is not self:
            mods = []
            for name, _, _, _ in self._table:
                mod = self
                for p in name.split(".")[:-1]:
                    mod = mod._modules[p]
                mods.append(mod)
            cache = (self, mods)
            self.__dict__["_leaf_cache"] = cache
        return cache[1]

    def _leaf_tensor(self, i):
        # parameters of a replica are plain attributes (no longer leaves), buffers stay in _buffers: getattr covers both
        return getattr(self._leaf_modules()[i], self._leaves[i][1])

    def _leaf_tensors(self):
        mods = self._leaf_modules()
        return [getattr(mods[i], self._leaves[i][1]) for i in range(len(mods))]

    def _replicate_for_data_parallel(self):
        """``nn.DataParallel.replicate`` support (the reference's ``DataParallelINNModel``, utils/parallel.py:150-169, calls
        replicas from one Python thread per GPU): a replica gets its OWN library handles, flat buffer and workspace --
        derived-weight caches are never shared between objects, because each object refreshes its handle from its own
        flat buffer."""
        replica = super()._replicate_for_data_parallel()
        d = replica.__dict__
        d["_handles"] = {}
        d["_flat"] = None
        d["_ws"] = {}
        d["_refreshed_for"] = None
        d["_param_epoch"] = None
        for k in ("_flat_param", "_train_mask", "_scratch_pool", "_leaf_cache", "_leaf_tuple", "_layout_ok_for", "flat_grad"):
            d.pop(k, None)
        return replica

    def _apply(self, fn, *a, **k):
        # .to() / .cuda() / .float() may replace the storage of every leaf: re-validate the flat layout on the next call
        self.__dict__["_layout_ok_for"] = None
        return super()._apply(fn, *a, **k)

    def _sync_flat(self, device):
        """All floating-point state lives in ONE flat buffer (the layout libtmglow_b200 reads);
        the module's parameters/buffers are views into it.  Rebuilt whenever ``.to()``/
        ``load_state_dict(assign=True)`` replaced the storage.  The full check (873 pointer comparisons, ~1.4 ms of
        host time) runs after an event that can move storage; otherwise two canary leaves are checked."""
        flat = self._flat
        ok = flat is not None and flat.device == device
        if ok and self.__dict__.get("_is_replica", False):
            return False                   # a replica lives for one scatterModel(): its flat copy was made on first use
        if ok and self.__dict__.get("_layout_ok_for") == (device, flat.data_ptr()):
            base, n = flat.data_ptr(), len(self._table)
            for i in (0, n // 2, n - 1):
                t = self._leaf_tensor(i)
                if t.data_ptr() != base + 4 * self._table[i][1] or t.dtype != torch.float32:
                    ok = False
                    break
            if ok:
                return False
            ok = True
        if ok:
            base = self._flat.data_ptr()
            for i, (name, off, numel, shape) in enumerate(self._table):
                t = self._leaf_tensor(i)
                if t.data_ptr() != base + 4 * off or t.dtype != torch.float32:
                    ok = False
                    break
        if ok:
            self.__dict__["_layout_ok_for"] = (device, self._flat.data_ptr())
            return False
        flat = torch.empty(self._n_flat, dtype=torch.float32, device=device)
        replica = bool(self.__dict__.get("_is_replica", False))
        with torch.no_grad():
            for i, (name, off, numel, shape) in enumerate(self._table):
                t = self._leaf_tensor(i)
                view = flat[off:off + numel].view(t.shape)
                view.copy_(t.detach().to(device=device, dtype=torch.float32))
                if not replica:
                    t.data = view          # a replica's tensors are broadcast copies owned by autograd: left alone
        self._flat = flat
        self._refreshed_for = None
        self.__dict__["_leaf_tuple"] = None
        self.__dict__["_layout_ok_for"] = (device, flat.data_ptr())
        return True

    def _version_sum(self):
        """Sum of the leaves' version counters (in-place writes by optimizers, ``load_state_dict``, ``copy_`` bump them; the
        sum only grows).  The tensor objects are cached while the flat layout stands."""
        tup = self.__dict__.get("_leaf_tuple")
        if tup is None or tup[0] is not self:
            tup = (self, tuple(self._leaf_tensors()))
            self.__dict__["_leaf_tuple"] = tup
        return sum(map(_VERSION, tup[1]))

    def _prepare(self, device):
        if device.type != "cuda":
            raise RuntimeError("tmglow_b200.TMGlow runs only on CUDA devices (no CPU fallback); got %s" % device)
        lib = _lib.load()
        changed = self._sync_flat(device)
        h = self._handle(device.index if device.index is not None else torch.cuda.current_device())
        # Derived weights (packed conv weights, W / W^-1 of the 1x1 convolutions, log-det constants) are re-derived
        # when the flat parameter buffer was replaced or any parameter/buffer was written in place since the last call
        # (optimizers, load_state_dict, copy_ bump the tensors' version counters; their sum only grows).
        # Writes through ``p.data`` bypass the counter: call ``refresh_weights()`` after those, or set
        # ``always_refresh_weights = True``.
        key = (h.value, self._flat.data_ptr(), self._version_sum())
        if changed or self.always_refresh_weights or self._refreshed_for != key:
            st = torch.cuda.current_stream(device).cuda_stream
            _lib.check(lib.tmg_model_refresh(h, self._flat.data_ptr(), st))
            self._refreshed_for = key
        if lib.tmg_model_get_precision(h) != _lib.PRECISIONS[self._precision]:
            _lib.check(lib.tmg_model_set_precision(h, _lib.PRECISIONS[self._precision]))
        return lib, h

    def refresh_weights(self):
        """Force the kernel-side weights to be re-derived on the next call (after writes through ``.data``)."""
        self._refreshed_for = None

    def _workspace(self, lib, h, B, hh, ww, device):
        key = (device, B, hh, ww)
        ws = self._ws.get(key)
        if ws is None:
            n = lib.tmg_workspace_bytes(h, B, hh, ww)
            if n == 0:
                raise AssertionError(lib.tmg_last_error().decode("utf-8", "replace"))
            self._ws.clear()          # keep one workspace alive
            ws = torch.empty(n, dtype=torch.uint8, device=device)
            self._ws[key] = ws
        return ws

    def _scratch(self, name, shape, dtype, device):
        """Persistent scratch tensors of the training backward (workspace, contiguous copies of the incoming gradients):
        used only inside one stream-ordered backward call, so one buffer per name is enough -- and stable addresses
        let the library replay its CUDA graph of the call instead of re-launching ~1 400 kernels."""
        pool = self.__dict__.setdefault("_scratch_pool", {})
        key = (name, tuple(shape), dtype, device)
        t = pool.get(key)
        if t is None:
            for k in [k for k in pool if k[0] == name]:
                del pool[k]
            t = torch.empty(shape, dtype=dtype, device=device)
            pool[key] = t
        return t

    # ------------------------------------------------------------------ shapes
    def _hf_size(self, x):
        up = self._cfg.cglow_upscale
        return x.shape[2] * up, x.shape[3] * up

    def latent_shapes(self, B, H, W):
        """Shapes of eps[0..L-1] (split noise) and eps[L] (top noise)."""
        shapes, c = [], self._cfg.out_features
        for _ in self.glow_blocks:
            c, H, W = c * 4, H // 2, W // 2
            shapes.append((B, c // 2, H, W))
            c //= 2
        shapes.append((B, c, H, W))
        return shapes

    def _state_dims(self, B, H, W):
        return [(B, self.rec_features, H >> (l + 1), W >> (l + 1)) for l in range(len(self.glow_blocks))]

    @staticmethod
    def _f32c(t, device):
        if t.device != device:
            raise RuntimeError("tensor on %s but the model runs on %s" % (t.device, device))
        return t.detach().to(torch.float32).contiguous()

    def _states_in(self, lib, h_in, dims, device, st, check_len=True):
        """LSTM states enter the library channels-last; NCHW tensors are converted by the
        library's own permutation kernel."""
        if h_in is None:
            return None, None, []
        assert not check_len or len(h_in) == len(self.glow_blocks), \
            'List of recurrent states need to be same length as flow blocks.'
        hs, cs, keep = [], [], []
        for (hh, cc), d in zip(h_in, dims):
            for t, out in ((hh, hs), (cc, cs)):
                assert tuple(t.shape) == d, "LSTM state shape %s, expected %s" % (tuple(t.shape), d)
                t = t.detach()
                if t.dtype != torch.float32:
                    t = t.float()
                if t.device != device:
                    raise RuntimeError("LSTM state on %s but the model runs on %s" % (t.device, device))
                same_memory = d[1] == 1 or d[2] * d[3] == 1      # NCHW and NHWC coincide
                if not (t.is_contiguous(memory_format=torch.channels_last) and not same_memory) and \
                        not (same_memory and t.is_contiguous()):
                    src = t.contiguous()
                    t = _empty_channels_last(d, device)
                    _lib.check(lib.tmg_nchw_to_nhwc(src.data_ptr(), t.data_ptr(), d[0], d[1], d[2], d[3], st))
                    keep.append(src)
                keep.append(t)
                out.append(t.data_ptr())
        return _lib.ptr_array(hs), _lib.ptr_array(cs), keep

    @staticmethod
    def _states_out(dims, device):
        hs = [_empty_channels_last(d, device) for d in dims]
        cs = [_empty_channels_last(d, device) for d in dims]
        return hs, cs

    def _flags(self, shared=False):
        return (_lib.TMG_FLAG_BN_TRAIN if self.training else 0) | (_lib.TMG_FLAG_SHARED_X if shared else 0)

    def _x_arg(self, x, device):
        """The LF input as the library wants it.  A batch-expanded view (``x1.expand(S, -1, -1, -1)``: one
        low-fidelity snapshot, S stochastic samples -- the uncertainty-quantification loop of
        trainFlowParallel.py:345-358 folded into the batch dimension) is passed as ONE input with
        TMG_FLAG_SHARED_X: the encoder runs once and the conditioning maps are shared by all samples.
        Results are those of the materialised batch.  Not used in train mode (BatchNorm batch statistics)."""
        shared = x.dim() == 4 and x.shape[0] > 1 and x.stride(0) == 0 and not self.training
        if shared:
            return self._f32c(x[:1], device), True
        return self._f32c(x, device), False

    def train(self, mode=True):
        # the eval-mode BatchNorm fold is derived from the running statistics, which train-mode calls update in place
        # inside the library: re-derive when the mode changes (not on every train-mode call)
        if mode != self.training:
            self._refreshed_for = None
        return super().train(mode)

    def _bump_bn_counters(self):
        if self.training:
            for name, b in self.named_buffers():
                if name.endswith("num_batches_tracked"):
                    b += 1

    # ------------------------------------------------------------------ public API (reference names)
    def forward(self, x, y, h_in=None, return_eps=False):
        """Encoder + flow y -> z with exact log-likelihood (reference nn/tmGlow.py:378-414).

        :returns: ``z [B,Cz,H/2^L,W/2^L]``, ``log_prior + log_det [B]``, ``h_out`` (list of (h, c)),
                  ``eps`` (list of L+1 noise tensors) or ``None``
        """
        device = x.device
        lib, h = self._prepare(device)
        with torch.cuda.device(device):
            st = torch.cuda.current_stream(device).cuda_stream
            B, H, W = x.shape[0], *self._hf_size(x)
            assert x.dim() == 4 and x.shape[1] == self._cfg.in_features, "x must be [B,%d,h,w]" % self._cfg.in_features
            x, shared = self._x_arg(x, device)
            y = self._f32c(y, device)
            assert tuple(y.shape) == (B, self._cfg.out_features, H, W), \
                "y must be [%d,%d,%d,%d], got %s" % (B, self._cfg.out_features, H, W, tuple(y.shape))
            ws = self._workspace(lib, h, B, x.shape[2], x.shape[3], device)
            dims = self._state_dims(B, H, W)
            hp, cp, keep = self._states_in(lib, h_in, dims, device, st)
            ho, co = self._states_out(dims, device)
            shapes = self.latent_shapes(B, H, W)
            z = torch.empty(shapes[-1], dtype=torch.float32, device=device)
            logp = torch.empty(B, dtype=torch.float32, device=device)
            eps = [torch.empty(s, dtype=torch.float32, device=device) for s in shapes] if return_eps else None
            _lib.check(lib.tmg_forward(
                h, B, x.shape[2], x.shape[3], x.data_ptr(), y.data_ptr(), hp, cp, z.data_ptr(), logp.data_ptr(),
                _lib.ptr_array([t.data_ptr() for t in ho]), _lib.ptr_array([t.data_ptr() for t in co]),
                _lib.ptr_array([t.data_ptr() for t in eps]) if return_eps else None,
                ws.data_ptr(), ws.numel(), self._flags(shared), st))
            self._bump_bn_counters()
        return z, logp, list(zip(ho, co)), eps

    def reconstruct(self, x, h_in, eps):
        """Encoder + flow z -> y with explicit noise (reference nn/tmGlow.py:442-467): ``eps[-1]`` is the
        top-latent noise, ``eps[:-1]`` the per-block split noise.

        :returns: ``y [B,out_features,H,W]``, ``log_det [B]`` (no top-prior term, like the reference), ``h_out``
        """
        device = x.device
        lib, h = self._prepare(device)
        with torch.cuda.device(device):
            st = torch.cuda.current_stream(device).cuda_stream
            assert x.dim() == 4 and x.shape[1] == self._cfg.in_features, "x must be [B,%d,h,w]" % self._cfg.in_features
            B, H, W = x.shape[0], *self._hf_size(x)
            x, shared = self._x_arg(x, device)
            shapes = self.latent_shapes(B, H, W)
            assert len(eps) == len(shapes), "eps must hold %d tensors" % len(shapes)
            eps_c = []
            for e, s in zip(eps, shapes):
                assert tuple(e.shape) == s, "eps shape %s, expected %s" % (tuple(e.shape), s)
                eps_c.append(self._f32c(e, device))
            ws = self._workspace(lib, h, B, x.shape[2], x.shape[3], device)
            dims = self._state_dims(B, H, W)
            hp, cp, keep = self._states_in(lib, h_in, dims, device, st)
            ho, co = self._states_out(dims, device)
            y = torch.empty((B, self._cfg.out_features, H, W), dtype=torch.float32, device=device)
            log_det = torch.empty(B, dtype=torch.float32, device=device)
            _lib.check(lib.tmg_reconstruct(
                h, B, x.shape[2], x.shape[3], x.data_ptr(), hp, cp,
                _lib.ptr_array([t.data_ptr() for t in eps_c]), y.data_ptr(), log_det.data_ptr(),
                _lib.ptr_array([t.data_ptr() for t in ho]), _lib.ptr_array([t.data_ptr() for t in co]),
                ws.data_ptr(), ws.numel(), self._flags(shared), st))
            self._bump_bn_counters()
        return y, log_det, list(zip(ho, co))

    # ------------------------------------------------------------------ training (reverse-KL training runs through sample())
    def flat_parameters(self):
        """The flat fp32 buffer that holds every parameter and buffer (the module's parameters are views of it).
        ``reconstruct_train`` returns gradients w.r.t. it; an optimizer can be built directly on it."""
        device = next(self.parameters()).device
        self._sync_flat(device)
        return self._flat

    def flat_parameter_for_optimizer(self):
        """An ``nn.Parameter`` aliasing the flat buffer (NOT registered on the module, so ``state_dict`` keeps the
        reference layout): build the optimizer on ``[model.flat_parameter_for_optimizer()]``.  Buffers inside the flat
        buffer (masks, running statistics) receive zero gradients, hence zero Adam updates; do not use weight decay."""
        flat = self.flat_parameters()
        fp = self.__dict__.get("_flat_param")
        if fp is None or fp.data_ptr() != flat.data_ptr():
            fp = torch.nn.Parameter(flat)
            object.__setattr__(self, "_flat_param", fp)
        return fp

    def reconstruct_train(self, x, h_in, eps):
        """Differentiable ``reconstruct`` (and therefore ``sample``): the forward runs in the model's precision mode
        and records the input of every flow step; ``backward`` (hand-written CUDA, exact fp32) returns the gradients
        w.r.t. the incoming LSTM states and ACCUMULATES the parameter gradients into ``self.flat_grad`` (a flat
        buffer laid out like ``flat_parameters()``; ``zero_flat_grad()`` clears it, ``scatter_flat_grad()`` exposes it
        as ``p.grad`` of every parameter).  All parameters are covered: the flow steps, split priors, ConvLSTM cells
        and the encoder (tests/test_gpu_backward.py checks every one against oracle autograd)."""
        # the flat parameter is passed so that the outputs carry a graph even without incoming states; its gradient
        # is accumulated into ``flat_grad`` by the backward kernel (autograd gets None for it)
        return _ReconstructFn.apply(self, x, eps, self.flat_parameter_for_optimizer(), *([t for hc in (h_in or []) for t in hc]))

    def sample_train(self, x, h_in=None):
        B, (H, W) = x.shape[0], self._hf_size(x)
        shapes = self.latent_shapes(B, H, W)
        eps = [None] * len(shapes)
        for i in [len(shapes) - 1] + list(range(len(shapes) - 2, -1, -1)):
            eps[i] = torch.randn(shapes[i], dtype=torch.float32, device=x.device)
        return self.reconstruct_train(x, h_in, eps)

    def reconstruct_block_train(self, x_block, h_in, eps_block):
        """A whole BPTT block (the ``tback`` chained ``sample()`` calls of nn/trainFlowParallel.py:248-277) as ONE
        differentiable call: ``x_block [B,T,nic,h,w]``, ``eps_block`` = list over the T time steps of the per-call noise
        lists (as ``reconstruct`` takes them).  Returns ``y [B,T,out,H,W]``, ``log_det [B,T]`` and the LSTM states after the
        last time step.  Same results as T chained ``reconstruct_train`` calls; inside the library the time steps are
        coupled only through the LSTM step of each level, everything else runs once on batch T*B (tmg_bptt_forward)."""
        return _BlockFn.apply(self, x_block, eps_block, self.flat_parameter_for_optimizer(), *([t for hc in (h_in or []) for t in hc]))

    def sample_block_train(self, x_block, h_in=None):
        """``reconstruct_block_train`` with the noise drawn like T successive ``sample()`` calls (same RNG consumption)."""
        B, T = x_block.shape[0], x_block.shape[1]
        H, W = x_block.shape[-2] * self._cfg.cglow_upscale, x_block.shape[-1] * self._cfg.cglow_upscale
        shapes = self.latent_shapes(B, H, W)
        eps_block = []
        for _ in range(T):
            eps = [None] * len(shapes)
            for i in [len(shapes) - 1] + list(range(len(shapes) - 2, -1, -1)):
                eps[i] = torch.randn(shapes[i], dtype=torch.float32, device=x_block.device)
            eps_block.append(eps)
        return self.reconstruct_block_train(x_block, h_in, eps_block)

    def zero_flat_grad(self):
        flat = self.flat_parameters()
        if getattr(self, "flat_grad", None) is None or self.flat_grad.shape != flat.shape or self.flat_grad.device != flat.device:
            self.flat_grad = torch.zeros_like(flat)
        else:
            self.flat_grad.zero_()
        return self.flat_grad

    def trainable_mask(self):
        """Flat 0/1 tensor laid out like ``flat_parameters()``: 1 on the entries that are ``nn.Parameter``s (the reference's
        ``model.parameters()``), 0 on buffers (masks, permutations, running statistics).  Used for weight decay."""
        flat = self.flat_parameters()
        mk = self.__dict__.get("_train_mask")
        if mk is None or mk.shape != flat.shape or mk.device != flat.device:
            mk = torch.zeros_like(flat)
            for i, (name, off, numel, shape) in enumerate(self._table):
                if self._leaves[i][2]:
                    mk[off:off + numel] = 1.0
            object.__setattr__(self, "_train_mask", mk)
        return mk

    def finalize_flat_grad(self):
        """Finish the gradients the per-time-step backward leaves in accumulated form (LU-parameterised 1x1 convolutions
        and their log-det terms: linear in quantities summed over the time steps of a BPTT block, so they are turned into
        the gradients of ``l, u, log_s`` and the ActNorm weights once, for all flow steps in one launch).  Call after the
        last ``backward()`` of an optimizer step and before reading ``flat_grad`` (``scatter_flat_grad`` and
        ``train.train_block`` do).  Idempotent: the accumulation slots are cleared."""
        g = getattr(self, "flat_grad", None)
        if g is None or g.device.type != "cuda":
            return g
        lib, h = self._prepare(g.device)
        with torch.cuda.device(g.device):
            _lib.check(lib.tmg_backward_finalize(h, g.data_ptr(), torch.cuda.current_stream(g.device).cuda_stream))
        return g

    def backward_graph_stats(self):
        """(graphs captured, graph replays, eager runs) of the per-time-step CUDA backward on the current device."""
        import ctypes as C
        device = next(self.parameters()).device
        lib, h = self._prepare(device)
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        _lib.check(lib.tmg_backward_graph_stats(h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def scatter_flat_grad(self):
        """``p.grad`` of every parameter = its slice of ``flat_grad`` (views, no copies)."""
        self.finalize_flat_grad()
        for i, (name, off, numel, shape) in enumerate(self._table):
            mod, attr, is_param = self._leaves[i]
            if is_param:
                mod._parameters[attr].grad = self.flat_grad[off:off + numel].view(shape)

    def sample(self, x, h_in=None):
        """Conditional generation (reference nn/tmGlow.py:417-440).  The Gaussian noise is drawn with
        ``torch.randn`` in the order and shapes of the reference (top latent first, then the splits of
        blocks L-1..0: tmGlow.py:435, flowUtils.py:206), so a seeded run consumes the RNG identically."""
        B, (H, W) = x.shape[0], self._hf_size(x)
        shapes = self.latent_shapes(B, H, W)
        eps = [None] * len(shapes)
        for i in [len(shapes) - 1] + list(range(len(shapes) - 2, -1, -1)):
            eps[i] = torch.randn(shapes[i], dtype=torch.float32, device=x.device)
        return self.reconstruct(x, h_in, eps)

    def encode(self, x):
        """``Encoder.forward`` (nn/tmGlow.py:104-129): returns ``(z_out, c_out)`` in NCHW."""
        device = x.device
        lib, h = self._prepare(device)
        with torch.cuda.device(device):
            st = torch.cuda.current_stream(device).cuda_stream
            x = self._f32c(x, device)
            B, H, W = x.shape[0], *self._hf_size(x)
            ws = self._workspace(lib, h, B, x.shape[2], x.shape[3], device)
            L = len(self.glow_blocks)
            c_out = [torch.empty((B, self._cfg.cond_features, H >> (l + 1), W >> (l + 1)), dtype=torch.float32,
                                 device=device) for l in range(L)]
            z_out = torch.empty((B, 2 * self.latent_shapes(B, H, W)[-1][1], H >> L, W >> L), dtype=torch.float32,
                                device=device)
            _lib.check(lib.tmg_encoder_forward(h, B, x.shape[2], x.shape[3], x.data_ptr(),
                                               _lib.ptr_array([t.data_ptr() for t in c_out]), z_out.data_ptr(),
                                               ws.data_ptr(), ws.numel(), self._flags(), st))
            self._bump_bn_counters()
        return z_out, c_out

    def activation_overflow(self, clear=True):
        """True when a kernel of the fp16-operand modes (``f16x3`` / ``f16``) had to clamp an activation to the fp16 range
        (|v| > 6e4) since the flag was last cleared -- the result then deviates from the reference's fp32 arithmetic
        (ill-conditioned weights; rerun in ``precision = "fp32"``).  Synchronises the device."""
        device = next(self.parameters()).device
        lib, h = self._prepare(device)
        with torch.cuda.device(device):
            r = lib.tmg_model_overflow(h, int(clear), torch.cuda.current_stream(device).cuda_stream)
        if r < 0:
            _lib.check(r)
        return r == 1

    def check_overflow(self):
        """Raises ``FloatingPointError`` (text of ``tmg_last_error``) when ``activation_overflow()``."""
        if self.activation_overflow(clear=True):
            raise FloatingPointError(_lib.load().tmg_last_error().decode("utf-8", "replace"))

    def _num_parameters(self):
        """Number of trainable parameters (reference nn/tmGlow.py:469-479)."""
        return sum(p.numel() for p in self.parameters())

    def initLSTMStates(self, seeds, input_dim):
        """Seeded initial LSTM states (reference nn/tmGlow.py:481-509): one CPU generator per sample,
        hidden state ~ U(-1,1), cell state ~ N(0,1), moved to the model's device."""
        device = next(self.parameters()).device
        states = []
        for i in range(len(self.glow_blocks)):
            dims = [1, self.rec_features, input_dim[0] // (2 ** (i + 1)), input_dim[1] // (2 ** (i + 1))]
            hs, cs = [], []
            for j in range(seeds.size(0)):
                gen = torch.Generator().manual_seed(int(seeds[j].item()))
                hs.append(2 * torch.rand(dims, generator=gen) - 1)
                cs.append(torch.randn(dims, generator=gen))
            states.append((torch.cat(hs, dim=0).to(device), torch.cat(cs, dim=0).to(device)))
        return states

    # ------------------------------------------------------------------ helpers used by tests
    def conv1x1_weight(self, level, step, inverse=False):
        """W (or W^-1) of ``InvertibleConv1x1LU`` as derived on the device (glowConv.py:151-174)."""
        device = next(self.parameters()).device
        lib, h = self._prepare(device)
        C = self._cfg.out_features * 4 * 2 ** level
        out = torch.empty((C, C), dtype=torch.float32, device=device)
        with torch.cuda.device(device):
            _lib.check(lib.tmg_model_get_conv1x1(h, level, step, int(inverse), out.data_ptr(),
                                                 torch.cuda.current_stream(device).cuda_stream))
        return out


class _BlockFn(torch.autograd.Function):
    """autograd bridge of a BPTT block: tmg_bptt_forward / tmg_bptt_backward (time-major tensors inside)."""

    @staticmethod
    def forward(ctx, model, x_block, eps_block, flat_param, *states):
        device = x_block.device
        lib, h = model._prepare(device)
        L = len(model.glow_blocks)
        with torch.cuda.device(device):
            st = torch.cuda.current_stream(device).cuda_stream
            B, T = x_block.shape[0], x_block.shape[1]
            x = x_block.detach().to(torch.float32).transpose(0, 1).contiguous()           # [T,B,nic,h,w]
            hh, ww = x.shape[-2], x.shape[-1]
            H, W = hh * model._cfg.cglow_upscale, ww * model._cfg.cglow_upscale
            shapes = model.latent_shapes(B, H, W)
            assert len(eps_block) == T and all(len(e) == len(shapes) for e in eps_block)
            eps_c = [torch.stack([model._f32c(eps_block[t][l], device) for t in range(T)], 0) for l in range(len(shapes))]
            for e, s in zip(eps_c, shapes):
                assert tuple(e.shape[1:]) == tuple(s)
            dims = model._state_dims(B, H, W)
            h_in = [(states[2 * l], states[2 * l + 1]) for l in range(L)] if states else None
            hp, cp, keep = model._states_in(lib, h_in, dims, device, st)
            ho, co = model._states_out(dims, device)
            y = torch.empty((T, B, model._cfg.out_features, H, W), dtype=torch.float32, device=device)
            log_det = torch.empty((T, B), dtype=torch.float32, device=device)
            n_ws = lib.tmg_bptt_workspace_bytes(h, T, B, hh, ww)
            if n_ws == 0:
                raise AssertionError(lib.tmg_last_error().decode("utf-8", "replace"))
            ws = model._scratch("bptt_ws", (n_ws,), torch.uint8, device)
            tape = torch.empty(lib.tmg_bptt_tape_bytes(h, T, B, hh, ww), dtype=torch.uint8, device=device)
            _lib.check(lib.tmg_bptt_forward(
                h, T, B, hh, ww, x.data_ptr(), hp, cp, _lib.ptr_array([t.data_ptr() for t in eps_c]),
                y.data_ptr(), log_det.data_ptr(), _lib.ptr_array([t.data_ptr() for t in ho]),
                _lib.ptr_array([t.data_ptr() for t in co]), tape.data_ptr(), tape.numel(), ws.data_ptr(), ws.numel(),
                model._flags(), st))
            if model.training:
                for name, b in model.named_buffers():
                    if name.endswith("num_batches_tracked"):
                        b += T
        ctx.model, ctx.x, ctx.eps, ctx.tape, ctx.keep, ctx.has_states = model, x, eps_c, tape, keep, bool(states)
        ctx.state_ptrs, ctx.dims, ctx.TB = (hp, cp), dims, (T, B)
        outs = [y.transpose(0, 1), log_det.transpose(0, 1)]
        for a, b in zip(ho, co):
            outs += [a, b]
        return tuple(outs)

    @staticmethod
    def backward(ctx, g_y, g_ld, *g_states):
        model, x = ctx.model, ctx.x
        device = x.device
        lib, h = model._prepare(device)
        L = len(model.glow_blocks)
        T, B = ctx.TB
        if getattr(model, "flat_grad", None) is None:
            model.zero_flat_grad()
        cl = lambda t: None if t is None else t.detach().float().contiguous(memory_format=torch.channels_last)
        with torch.cuda.device(device):
            st = torch.cuda.current_stream(device).cuda_stream
            H, W = x.shape[-2] * model._cfg.cglow_upscale, x.shape[-1] * model._cfg.cglow_upscale
            gy_buf = model._scratch("g_yb", (T, B, model._cfg.out_features, H, W), torch.float32, device)
            gld_buf = model._scratch("g_ldb", (T, B), torch.float32, device)
            if g_y is None:
                gy_buf.zero_()
            else:
                gy_buf.copy_(g_y.transpose(0, 1))
            if g_ld is None:
                gld_buf.zero_()
            else:
                gld_buf.copy_(g_ld.transpose(0, 1))
            gh = [cl(g_states[2 * l]) for l in range(L)]
            gc = [cl(g_states[2 * l + 1]) for l in range(L)]
            g_in = [(_empty_channels_last(d, device), _empty_channels_last(d, device)) for d in ctx.dims] if ctx.has_states else []
            ws = model._scratch("bptt_ws", (lib.tmg_bptt_workspace_bytes(h, T, B, x.shape[-2], x.shape[-1]),), torch.uint8, device)
            hp, cp = ctx.state_ptrs
            pa = lambda ts: _lib.ptr_array([None if t is None else t.data_ptr() for t in ts])
            _lib.check(lib.tmg_bptt_backward(
                h, T, B, x.shape[-2], x.shape[-1], x.data_ptr(), hp, cp, _lib.ptr_array([t.data_ptr() for t in ctx.eps]),
                ctx.tape.data_ptr(), gy_buf.data_ptr(), gld_buf.data_ptr(), pa(gh), pa(gc),
                pa([a for a, _ in g_in]) if g_in else None, pa([b for _, b in g_in]) if g_in else None,
                model.flat_grad.data_ptr(), ws.data_ptr(), ws.numel(), model._flags(), st))
        grads = [None, None, None, None]
        for a, b in g_in:
            grads += [a, b]
        return tuple(grads)


class _ReconstructFn(torch.autograd.Function):
    """autograd bridge of ``TMGlow.reconstruct``: tmg_reconstruct_train / tmg_reconstruct_backward."""

    @staticmethod
    def forward(ctx, model, x, eps, flat_param, *states):
        device = x.device
        lib, h = model._prepare(device)
        L = len(model.glow_blocks)
        with torch.cuda.device(device):
            st = torch.cuda.current_stream(device).cuda_stream
            x = model._f32c(x, device)
            B, H, W = x.shape[0], *model._hf_size(x)
            shapes = model.latent_shapes(B, H, W)
            eps_c = [model._f32c(e, device) for e in eps]
            assert [tuple(e.shape) for e in eps_c] == [tuple(s) for s in shapes]
            dims = model._state_dims(B, H, W)
            h_in = [(states[2 * l], states[2 * l + 1]) for l in range(L)] if states else None
            hp, cp, keep = model._states_in(lib, h_in, dims, device, st)
            ho, co = model._states_out(dims, device)
            y = torch.empty((B, model._cfg.out_features, H, W), dtype=torch.float32, device=device)
            log_det = torch.empty(B, dtype=torch.float32, device=device)
            ws = model._workspace(lib, h, B, x.shape[2], x.shape[3], device)
            tape = torch.empty(lib.tmg_tape_bytes(h, B, x.shape[2], x.shape[3]), dtype=torch.uint8, device=device)
            _lib.check(lib.tmg_reconstruct_train(
                h, B, x.shape[2], x.shape[3], x.data_ptr(), hp, cp, _lib.ptr_array([t.data_ptr() for t in eps_c]),
                y.data_ptr(), log_det.data_ptr(), _lib.ptr_array([t.data_ptr() for t in ho]),
                _lib.ptr_array([t.data_ptr() for t in co]), tape.data_ptr(), tape.numel(), ws.data_ptr(), ws.numel(),
                model._flags(), st))
            model._bump_bn_counters()
        ctx.model, ctx.x, ctx.eps, ctx.tape, ctx.keep, ctx.has_states = model, x, eps_c, tape, keep, bool(states)
        ctx.state_ptrs = (hp, cp)
        ctx.dims = dims
        outs = [y, log_det]
        for a, b in zip(ho, co):
            outs += [a, b]
        return tuple(outs)

    @staticmethod
    def backward(ctx, g_y, g_ld, *g_states):
        model, x = ctx.model, ctx.x
        device = x.device
        lib, h = model._prepare(device)
        L = len(model.glow_blocks)
        B = x.shape[0]
        if getattr(model, "flat_grad", None) is None:
            model.zero_flat_grad()
        cl = lambda t, d: None if t is None else t.detach().float().contiguous(memory_format=torch.channels_last)
        with torch.cuda.device(device):
            st = torch.cuda.current_stream(device).cuda_stream
            gy_buf = model._scratch("g_y", (B, model._cfg.out_features) + tuple(model._hf_size(x)), torch.float32, device)
            gld_buf = model._scratch("g_ld", (B,), torch.float32, device)
            if g_y is None:
                gy_buf.zero_()
            else:
                gy_buf.copy_(g_y)
            if g_ld is None:
                gld_buf.zero_()
            else:
                gld_buf.copy_(g_ld)
            g_y, g_ld = gy_buf, gld_buf
            gh = [cl(g_states[2 * l], None) for l in range(L)]
            gc = [cl(g_states[2 * l + 1], None) for l in range(L)]
            g_in = [(_empty_channels_last(d, device), _empty_channels_last(d, device)) for d in ctx.dims] if ctx.has_states else []
            n = lib.tmg_reconstruct_backward_workspace_bytes(h, B, x.shape[2], x.shape[3])
            ws = model._scratch("bwd_ws", (n,), torch.uint8, device)
            hp, cp = ctx.state_ptrs
            pa = lambda ts: _lib.ptr_array([None if t is None else t.data_ptr() for t in ts])
            _lib.check(lib.tmg_reconstruct_backward(
                h, B, x.shape[2], x.shape[3], x.data_ptr(), hp, cp, _lib.ptr_array([t.data_ptr() for t in ctx.eps]),
                ctx.tape.data_ptr(), g_y.data_ptr(), g_ld.data_ptr(), pa(gh), pa(gc),
                pa([a for a, _ in g_in]) if g_in else None, pa([b for _, b in g_in]) if g_in else None,
                model.flat_grad.data_ptr(), ws.data_ptr(), ws.numel(), model._flags(), st))
        grads = [None, None, None, None]
        for a, b in g_in:
            grads += [a, b]
        return tuple(grads)
