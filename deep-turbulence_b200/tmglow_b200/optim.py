"""``FlatAdam``: the reference trainer's optimizer step -- ``clip_grad_norm_`` + ``Adam(weight_decay=1e-8, amsgrad=True).step()``
(``nn/trainFlowParallel.py:290-291``, ``main.py:78``) -- on the model's flat parameter as ONE library call (two kernels,
``csrc/optim.cu``) with no host synchronisation: the gradient norm, the clip coefficient, the step counter and the
hyper-parameters stay in device memory, so a training step can be captured in a CUDA graph and the Python thread never waits.

It IS a ``torch.optim.Adam`` (state layout, ``state_dict`` / ``load_state_dict``, ``param_groups``, lr schedulers), so
``workspace.saveWorkspace`` / ``load_flat_optimizer_state`` convert to and from the reference's per-parameter checkpoints
unchanged.  Only ``step`` differs: ``fused_step(grad, max_norm, weight_decay, mask)``.
"""
import torch

from . import _lib


class FlatAdam(torch.optim.Adam):
    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, amsgrad=True):
        self._model = model
        fp = model.flat_parameter_for_optimizer()
        # weight_decay = 0 inside torch's bookkeeping: the decay is applied to the trainable entries only (fused_step)
        super().__init__([fp], lr=lr, betas=betas, eps=eps, weight_decay=0.0, amsgrad=amsgrad)
        self._hyper = None
        self._hyper_host = None
        self._out = None
        self._ws = None
        self.reference_weight_decay = 0.0

    # ------------------------------------------------------------------ device-side mirrors
    def _param(self):
        return self.param_groups[0]["params"][0]

    def _ensure_state(self):
        p = self._param()
        st = self.state[p]
        if len(st) == 0:
            st["step"] = torch.tensor(0.0, dtype=torch.float32)
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            if self.param_groups[0]["amsgrad"]:
                st["max_exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    def _sync_hyper(self, max_norm, weight_decay):
        g = self.param_groups[0]
        st = self._ensure_state()
        dev = self._param().device
        host = (float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]), float(weight_decay or 0.0),
                float(max_norm) if max_norm else 0.0, float(st["step"]), 1.0 if g["amsgrad"] else 0.0)
        if self._hyper is None or self._hyper.device != dev:
            self._hyper = torch.tensor(host, dtype=torch.float32, device=dev)
            self._out = torch.zeros(2, dtype=torch.float32, device=dev)
            self._ws = torch.empty(_lib.load().tmg_adam_workspace_bytes(), dtype=torch.uint8, device=dev)
        elif self._hyper_host is None or host[:6] + host[7:] != self._hyper_host[:6] + self._hyper_host[7:] or \
                host[6] != self._hyper_host[6] + 1.0:
            # a hyper-parameter changed (lr schedule, load_state_dict): rewrite the device copy; in the steady state the kernel
            # itself advances the step counter and nothing is copied
            self._hyper.copy_(torch.tensor(host, dtype=torch.float32), non_blocking=True)
        self._hyper_host = host

    # ------------------------------------------------------------------ the step
    @torch.no_grad()
    def fused_step(self, grad, max_norm=None, weight_decay=0.0, mask=None):
        """Clip ``grad`` (flat, like the parameter) to ``max_norm``, add ``weight_decay * p`` on the entries where ``mask``
        is 1, Adam / AMSGrad update.  Returns a device tensor ``[norm_before_clipping, clip_coefficient]`` (no sync)."""
        p = self._param()
        if p.device.type != "cuda":
            raise RuntimeError("FlatAdam.fused_step runs only on CUDA devices (no CPU fallback)")
        st = self._ensure_state()
        self._sync_hyper(max_norm, weight_decay)
        lib = _lib.load()
        with torch.cuda.device(p.device):
            _lib.check(lib.tmg_adam_step(
                p.data_ptr(), grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                st["max_exp_avg_sq"].data_ptr() if "max_exp_avg_sq" in st else None,
                mask.data_ptr() if mask is not None else None, p.numel(), self._hyper.data_ptr(), self._out.data_ptr(),
                self._ws.data_ptr(), self._ws.numel(), torch.cuda.current_stream(p.device).cuda_stream))
        st["step"] += 1.0                    # host mirror of the device counter (checkpoints, bias-correction bookkeeping)
        self._model.refresh_weights()         # the write went through the library: derived (packed) weights follow
        return self._out

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._hyper_host = None              # the step counter / hyper-parameters may have changed: rewrite the device copy
