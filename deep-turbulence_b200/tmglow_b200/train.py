"""Data-parallel training step on top of ``TMGlow.sample_train`` (hand-written CUDA forward + backward).

Mirrors the inner loop of the reference trainer (``nn/trainFlowParallel.py:241-303``): reverse-KL training runs through
``sample()`` -- back-propagation through time over a block of ``tback`` time steps with the ConvLSTM states carried
from step to step --, the gradient norm is clipped, the optimizer steps once per block and the states are detached.
The reference replicates the model with ``nn.DataParallel`` threads inside one process (``utils/parallel.py``); here
every rank is one process with one GPU and the ONLY collective is one all-reduce of the flat gradient buffer per
optimizer step (all parameters live in one flat buffer, so there is nothing to bucket).  Works on NCCL (GPU) and on
gloo (the averaging / clipping logic is tested at world size 2 on CPU, tests/test_parallel_cpu.py).
"""
import math
from typing import Callable, List, Optional, Sequence, Tuple

import torch


def allreduce_mean_(t: torch.Tensor, group=None) -> torch.Tensor:
    """In-place mean over the ranks of ``group`` (no-op without an initialised process group).  Equal shards:
    mean over ranks of per-rank means == the reference's mean over GPUs (trainFlowParallel.py:285)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, group=group)
        t.div_(dist.get_world_size(group))
    return t


def clip_flat_grad_(g: torch.Tensor, max_norm: float) -> float:
    """``torch.nn.utils.clip_grad_norm_`` on the flat gradient (trainFlowParallel.py:290): returns the norm before."""
    norm = float(g.norm())
    if max_norm is not None and norm > max_norm:
        g.mul_(max_norm / (norm + 1e-6))
    return norm


def reverse_kl_loss(y_pred: torch.Tensor, log_det: torch.Tensor, target: torch.Tensor, beta: float = 200.0) -> torch.Tensor:
    """Data terms + entropy term only (torch ops; kept for tests of the BPTT chain).  The full reference loss with the
    PDE-residual terms is ``tmglow_b200.loss.TMGLowLoss`` (one fused CUDA kernel) -- pass it to ``train_block`` as
    ``criterion``.  y_pred, target: [B,T,C,H,W]; log_det: [B,T]."""
    mse = torch.mean((y_pred - target) ** 2)
    pred_rms = torch.sqrt(torch.mean((y_pred - y_pred.mean(dim=1, keepdim=True)) ** 2, dim=1) + 1e-12)
    tgt_rms = torch.sqrt(torch.mean((target - target.mean(dim=1, keepdim=True)) ** 2, dim=1) + 1e-12)
    rms = torch.mean((pred_rms - tgt_rms) ** 2)
    n_out = y_pred.shape[-3] * y_pred.shape[-2] * y_pred.shape[-1]
    neg_entropy = log_det.mean() / math.log(2.0) / n_out
    return beta * (mse + rms) + neg_entropy


def train_block(model, optimizer, x_block: torch.Tensor, target: torch.Tensor, h_in: Optional[list],
                loss_fn: Callable = reverse_kl_loss, max_norm: Optional[float] = 1.0, group=None,
                criterion=None, target_mean: Optional[torch.Tensor] = None, target_rms: Optional[torch.Tensor] = None,
                weight_decay: float = 0.0, block_call: bool = True):
    """One optimizer step on a BPTT block (trainFlowParallel.py:248-293).  ``x_block [B,T,nic,h,w]``, ``target
    [B,T,noc,H,W]``, ``h_in`` list of (h, c) or None; ``optimizer`` must have been built on
    ``[model.flat_parameter_for_optimizer()]``.  With ``criterion`` (a ``TMGLowLoss``) the loss is the reference's
    ``criterion(yPred, logp, target, target_mean, target_rms)`` with the statistics of the full series
    (``loss.target_statistics``); otherwise ``loss_fn(yPred, logp, target)``.
    With a ``tmglow_b200.optim.FlatAdam`` optimizer the clip / decay / AMSGrad update is one fused library call without a
    host synchronisation and ``grad_norm`` is returned as a 0-dim device tensor; with a plain ``torch.optim.Adam`` it is a float.
    ``block_call``: evaluate the block through ``model.sample_block_train`` (one library call, time-batched); ``False`` runs
    the ``T`` chained ``sample_train`` calls of round 1 (same results, ~5x the launches).
    Returns ``(loss, grad_norm, h_out)`` with ``h_out`` detached (truncated BPTT, trainFlowParallel.py:296-300)."""
    T = x_block.shape[1]
    model.zero_flat_grad()
    if block_call and hasattr(model, "sample_block_train"):
        # the whole block in one library call: time steps are coupled only through the LSTM step of each level
        outs = model.sample_block_train(x_block, h_in)
        y_pred, logp = outs[0], outs[1]
        h = [(outs[2 + 2 * l], outs[3 + 2 * l]) for l in range(len(model.glow_blocks))]
    else:
        ys, lds = [], []
        h = h_in
        for t in range(T):
            outs = model.sample_train(x_block[:, t], h)
            ys.append(outs[0]); lds.append(outs[1])
            h = [(outs[2 + 2 * l], outs[3 + 2 * l]) for l in range(len(model.glow_blocks))]
        y_pred, logp = torch.stack(ys, 1), torch.stack(lds, 1)
    if criterion is not None:
        loss = criterion(y_pred, logp, target, target_mean, target_rms)
    else:
        loss = loss_fn(y_pred, logp, target)
    loss.backward()
    g = model.flat_grad
    if hasattr(model, "finalize_flat_grad"):
        model.finalize_flat_grad()                  # deferred LU-parameter gradients, once per optimizer step
    allreduce_mean_(g, group)                       # the one collective of data-parallel training
    if hasattr(optimizer, "fused_step"):
        # clip + weight decay (trainable entries) + AMSGrad in one library call; the norm stays on the device (no sync)
        out = optimizer.fused_step(g, max_norm=max_norm, weight_decay=weight_decay, mask=model.trainable_mask())
        return loss.detach(), out[0], [(a.detach(), b.detach()) for a, b in h]
    norm = clip_flat_grad_(g, max_norm)
    flat = model.flat_parameter_for_optimizer()
    if weight_decay:
        # torch.optim.Adam(weight_decay=wd) adds wd * p to the (clipped) gradient inside step() (main.py:78: 1e-8); done
        # here on the trainable entries only -- the flat buffer also holds masks / permutations / running statistics,
        # which must see a zero gradient (build the optimizer WITHOUT weight_decay)
        g.addcmul_(flat.detach(), model.trainable_mask(), value=weight_decay)
    flat.grad = g
    optimizer.step()
    model.refresh_weights()                         # derived (packed) weights follow the new parameters
    return loss.detach(), norm, [(a.detach(), b.detach()) for a, b in h]


def mix_states(h_out: list, h_key: list) -> list:
    """trainFlowParallel.py:296-300: after each optimizer step the carried LSTM states are averaged with the initial
    ("key") states of the series: ``0.5*a_out.detach() + 0.5*a_key``."""
    return [(0.5 * a.detach() + 0.5 * ak, 0.5 * c.detach() + 0.5 * ck) for (a, c), (ak, ck) in zip(h_out, h_key)]


def train_series(model, optimizer, criterion, input0: torch.Tensor, target0: torch.Tensor, lstm_seeds: torch.Tensor,
                 tback: int = 10, max_norm: Optional[float] = 1.0, group=None, weight_decay: float = 1e-8):
    """One mini-batch of ``TrainFlow.trainParallel`` (trainFlowParallel.py:225-303): a time series ``input0
    [B,Tmax,nic,h,w]`` / ``target0 [B,Tmax,noc,H,W]`` is cut into ``Tmax // tback`` BPTT blocks, one optimizer step
    each; LSTM states start from ``initLSTMStates(lstm_seeds)`` and are mixed with them between blocks.
    ``weight_decay`` is the reference's Adam weight decay (1e-8, main.py:78), applied by ``train_block`` to the trainable
    entries of the flat buffer only.  Returns the summed loss (``total_loss``) and the final states."""
    dev = model.flat_parameter_for_optimizer().device
    a_key = [(a.to(dev), c.to(dev)) for a, c in model.initLSTMStates(lstm_seeds, [target0.size(-2), target0.size(-1)])]
    a0 = a_key
    target0 = target0.to(dev)
    from .loss import target_statistics
    t_mean, t_rms = target_statistics(target0)
    total = torch.zeros((), device=dev)
    for i in range(target0.size(1) // tback):
        xb = input0[:, i * tback:(i + 1) * tback].to(dev, non_blocking=True)
        tb = target0[:, i * tback:(i + 1) * tback]
        loss, _, a_out = train_block(model, optimizer, xb, tb, a0, max_norm=max_norm, group=group, criterion=criterion,
                                     target_mean=t_mean, target_rms=t_rms, weight_decay=weight_decay)
        a0 = mix_states(a_out, a_key)
        total = total + loss
    return total, a0


class GraphedTrainBlock:
    """One optimizer step of the reference trainer's inner loop (``nn/trainFlowParallel.py:248-303``: BPTT block through
    ``sample()``, ``TMGLowLoss``, backward, gradient all-reduce, clip, Adam-amsgrad, LSTM-state mixing) captured ONCE as CUDA
    graphs and replayed: the ~4 000 kernel launches of a step become two ``cudaGraphLaunch`` calls and the Python thread does
    no per-launch work -- what limits the step when the per-GPU batch is small (strong scaling).

    Static buffers: ``x`` [B,T,nic,h,w], ``target`` [B,T,noc,H,W], ``target_mean`` / ``target_rms`` [B,noc,H,W] and the LSTM
    states; ``step(x, target, ...)`` copies into them (device-to-device or pinned host-to-device, asynchronous) and replays.
    Graph A = zero the gradient, forward block, loss, backward, deferred LU gradients; then the gradient all-reduce (eager NCCL,
    only when a process group is active); graph B = ``FlatAdam.fused_step`` + state mixing with the key states.  The noise is
    drawn inside graph A by ``torch.randn`` (graph-safe CUDA generator: every replay consumes fresh offsets).
    Needs a ``FlatAdam`` optimizer (no host synchronisation inside the step)."""

    def __init__(self, model, optimizer, criterion, x_block, target, target_mean, target_rms, h_key, max_norm=1.0,
                 weight_decay=1e-8, group=None, warmup=2):
        assert hasattr(optimizer, "fused_step"), "GraphedTrainBlock needs tmglow_b200.optim.FlatAdam"
        self.model, self.opt, self.crit, self.group = model, optimizer, criterion, group
        self.max_norm, self.wd = max_norm, weight_decay
        dev = x_block.device
        self.x = x_block.clone(); self.target = target.clone()
        self.t_mean = target_mean.clone(); self.t_rms = target_rms.clone()
        self.h_key = [(a.clone(), c.clone()) for a, c in h_key]
        self.h_in = [(a.clone(), c.clone()) for a, c in h_key]
        self.loss = torch.zeros((), device=dev)
        self.norm = None
        self.replays = 0
        import torch.distributed as dist
        self._dist = dist if (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1) else None
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 2)):          # allocator, derived weights, optimizer state, kernel attributes
                self._part_a()
                self._allreduce()
                self._part_b()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        from . import _lib
        lib = _lib.load()
        n0 = lib.tmg_launch_count(0)
        self.ga, self.gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.ga):
            self._part_a()
        with torch.cuda.graph(self.gb, pool=self.ga.pool()):
            self._part_b()
        self.kernels_per_step = int(lib.tmg_launch_count(0) - n0)     # library kernels inside the two graphs (torch's not counted)
        self.time_allreduce = False
        self._ar_events = []

    def _part_a(self):
        m = self.model
        m.zero_flat_grad()
        outs = m.sample_block_train(self.x, self.h_in)
        L = len(m.glow_blocks)
        self._h_out = [(outs[2 + 2 * l].detach(), outs[3 + 2 * l].detach()) for l in range(L)]
        loss = self.crit(outs[0], outs[1], self.target, self.t_mean, self.t_rms)
        loss.backward()
        m.finalize_flat_grad()
        self.loss.copy_(loss.detach())

    def _allreduce(self):
        if self._dist is not None:
            if getattr(self, "time_allreduce", False):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                allreduce_mean_(self.model.flat_grad, self.group)
                e1.record()
                self._ar_events.append((e0, e1))
            else:
                allreduce_mean_(self.model.flat_grad, self.group)

    def allreduce_ms(self):
        """Device time of the gradient all-reduces recorded since ``time_allreduce`` was switched on (synchronises)."""
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in self._ar_events)
        n = len(self._ar_events)
        self._ar_events = []
        return ms, n

    def _part_b(self):
        out = self.opt.fused_step(self.model.flat_grad, max_norm=self.max_norm, weight_decay=self.wd, mask=self.model.trainable_mask())
        self.norm = out
        for (hi, ci), (ho, co), (hk, ck) in zip(self.h_in, self._h_out, self.h_key):      # trainFlowParallel.py:296-300
            hi.copy_(0.5 * ho + 0.5 * hk)
            ci.copy_(0.5 * co + 0.5 * ck)

    def load(self, x_block=None, target=None, target_mean=None, target_rms=None, h_key=None):
        """Asynchronous copies into the static buffers (pinned host tensors or device tensors)."""
        for dst, src in ((self.x, x_block), (self.target, target), (self.t_mean, target_mean), (self.t_rms, target_rms)):
            if src is not None:
                dst.copy_(src, non_blocking=True)
        if h_key is not None:                         # a new series: states restart from its key states
            for (hk, ck), (hi, ci), (a, c) in zip(self.h_key, self.h_in, h_key):
                hk.copy_(a, non_blocking=True); ck.copy_(c, non_blocking=True)
                hi.copy_(hk); ci.copy_(ck)

    def step(self):
        """Replays the captured step; returns ``(loss, grad_norm)`` as device tensors (no synchronisation)."""
        self.ga.replay()
        self._allreduce()
        self.gb.replay()
        self.replays += 1
        st = self.opt.state[self.opt.param_groups[0]["params"][0]]
        st["step"] += 1.0                             # host mirror of the device-side step counter
        if self.opt._hyper_host is not None:
            hh = list(self.opt._hyper_host); hh[6] += 1.0; self.opt._hyper_host = tuple(hh)
        self.model.refresh_weights()                  # the parameters changed inside the graph
        return self.loss, self.norm[0]
