"""Sample-parallel uncertainty quantification on top of ``TMGlow.sample``.

The reference draws stochastic high-fidelity samples of one low-fidelity sequence with a Python loop over
samples around the loop over time steps (``utils/utils.py:197-222`` ``modelPred``, ``nn/trainFlowParallel.py:345-367``
``TrainFlow.test``), one ``model.sample`` call per (sample, time step).  Here the sample loop is folded into the batch
dimension: one call per time step produces all S samples of this rank from ONE low-fidelity snapshot (passed as a
batch-expanded view, so the encoder runs once and the conditioning maps are shared), the ConvLSTM states stay on the
device, and only per-time-step moments (mean / variance over samples) leave the step.

Multi-GPU: samples are independent (own noise, own LSTM state), so the S samples are sharded
over the ranks with no data-path collective; the only communication is one all-reduce of the moment sums per
sequence (``combine_moments``), which also works on the gloo backend (tests/test_parallel_cpu.py).
"""
from typing import Callable, List, Optional, Sequence, Tuple

import torch


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced partition of ``range(total)``: ranks below ``total % world`` get one extra item."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world %d/%d" % (rank, world))
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def sample_seeds(base_seed: int, sequence: int, lo: int, hi: int) -> torch.Tensor:
    """LSTM-state seeds of samples [lo, hi) of one LF sequence: a function of the GLOBAL sample index only, so a run
    sharded over any number of ranks draws the states a single-rank run draws (``initLSTMStates`` seeds one CPU
    generator per sample, tmGlow.py:481-509)."""
    idx = torch.arange(lo, hi, dtype=torch.long)
    return (base_seed + 1000003 * sequence + idx) % (2 ** 31 - 1)


def combine_moments(s1: torch.Tensor, s2: torch.Tensor, n: int, group=None):
    """All-reduce the per-rank sums, sums of squares and counts -> (mean, unbiased variance, total count)."""
    import torch.distributed as dist
    cnt = torch.tensor([float(n)], dtype=torch.float64, device=s1.device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        s1 = s1.clone(); s2 = s2.clone()
        dist.all_reduce(s1, group=group)
        dist.all_reduce(s2, group=group)
        dist.all_reduce(cnt, group=group)
    ntot = int(cnt.item())
    mean = s1 / ntot
    var = (s2 - ntot * mean * mean) / max(ntot - 1, 1)
    return mean, var.clamp_(min=0), ntot


def mix_states(states, key, weight=0.5):
    """Average the running LSTM states with the initial ("key") states, as the reference does every 10 (test) or 20
    (modelPred) time steps: trainFlowParallel.py:360-366, utils.py:216-222."""
    return [(weight * h + (1 - weight) * hk, weight * c + (1 - weight) * ck) for (h, c), (hk, ck) in zip(states, key)]


@torch.no_grad()
def sample_sequence(model, x_seq: torch.Tensor, samples: int, *, base_seed: int = 0, sequence: int = 0,
                    state_mix_every: int = 10, rank: int = 0, world: int = 1, group=None,
                    unnormalise: bool = True, keep_samples: bool = False,
                    sampler: Optional[Callable] = None, init_states: Optional[Callable] = None,
                    key_states: Optional[list] = None):
    """S stochastic HF samples of one LF sequence ``x_seq [T, nic, h, w]``.

    ``key_states``: this rank's initial LSTM states when the caller prepared them ahead of time (``initLSTMStates`` is the
    reference's seeded HOST generator, one per sample: seconds of CPU time for thousands of samples); default: drawn here
    from ``sample_seeds`` of the rank's global sample indices.

    Returns ``(mean [T,C,H,W], var [T,C,H,W], n_total, samples or None)``; ``samples`` ([S_rank,T,C,H,W], this rank's
    shard) only when ``keep_samples``.  ``sampler(x, h) -> (y, log_det, h)`` and ``init_states(seeds, [H,W])`` default
    to ``model.sample`` / ``model.initLSTMStates`` (tests substitute CPU stand-ins).
    """
    sampler = sampler or model.sample
    init_states = init_states or model.initLSTMStates
    T = x_seq.shape[0]
    # checked on EVERY rank before any work: an empty shard would leave the other ranks waiting in the all-reduce
    if samples < world:
        raise ValueError("samples (%d) < world (%d): every rank needs at least one sample" % (samples, world))
    lo, hi = shard_range(samples, rank, world)
    S = hi - lo
    dev = x_seq.device
    out_mu = out_std = None
    if unnormalise and getattr(model, "out_mu", None) is not None:
        out_mu = model.out_mu.to(dev).view(1, -1, 1, 1)
        out_std = model.out_std.to(dev).view(1, -1, 1, 1)
        if float(out_std.abs().sum()) == 0.0:      # buffers never set by a data loader (dataLoader.py:159-164)
            out_mu = out_std = None
    s1 = s2 = None
    kept: List[torch.Tensor] = []
    if S > 0:
        up = getattr(getattr(model, "_cfg", None), "cglow_upscale", None)
        H, W = (x_seq.shape[-2] * up, x_seq.shape[-1] * up) if up else (None, None)
        key = key_states if key_states is not None else init_states(sample_seeds(base_seed, sequence, lo, hi), [H, W])
        h = key
    for t in range(T):
        if S > 0:
            x = x_seq[t:t + 1].expand(S, -1, -1, -1)          # ONE input, S samples: shared-input fast path
            y, _, h = sampler(x, h)
            if out_mu is not None:
                y = out_std * y + out_mu
            y64 = y.double()
            m1, m2 = y64.sum(0), (y64 * y64).sum(0)
            if keep_samples:
                kept.append(y)
            if state_mix_every and t % state_mix_every == 0:
                h = mix_states(h, key)
        else:
            m1 = m2 = None
        if s1 is None and m1 is not None:
            s1 = torch.zeros((T,) + tuple(m1.shape), dtype=torch.float64, device=dev)
            s2 = torch.zeros_like(s1)
        if m1 is not None:
            s1[t], s2[t] = m1, m2
    assert s1 is not None      # samples >= world: every rank has a shard
    mean, var, ntot = combine_moments(s1, s2, S, group)
    return mean.float(), var.float(), ntot, (torch.stack(kept, 1) if keep_samples else None)


@torch.no_grad()
def model_pred(model, input_seq: torch.Tensor, samples: int, tmax: int, *, stride: int = 1, state_mix_every: int = 20,
               seeds: Optional[torch.Tensor] = None, unnormalise: bool = True,
               sampler: Optional[Callable] = None, init_states: Optional[Callable] = None):
    """The prediction loops of the reference, ``modelPred`` (``utils/utils.py:151-235``, state mixing every 20 steps) and
    the body of ``TrainFlow.test`` (``nn/trainFlowParallel.py:345-367``, every 10 steps), for a mini-batch of ``B`` test
    cases ``input_seq [B, T, nic, h, w]``: ``samples`` stochastic predictions of ``tmax`` time steps each.  The reference
    runs ``samples x tmax`` calls of ``model.sample`` on batch ``B``; here the sample loop is folded into the batch
    (one call per time step on batch ``samples * B``, sample-major), each (sample, case) pair with its own seeded LSTM
    state.  ``seeds [samples, B]`` (default: drawn like the reference, ``LongTensor.random_(0, 1e8)``).
    Returns ``yPred [samples, B, tmax // stride, C, H, W]`` (un-normalised with ``model.out_mu / out_std`` when set)."""
    sampler = sampler or model.sample
    init_states = init_states or model.initLSTMStates
    B = input_seq.shape[0]
    dev = input_seq.device
    if seeds is None:
        seeds = torch.LongTensor(samples, B).random_(0, int(1e8))
    assert tuple(seeds.shape) == (samples, B)
    up = getattr(getattr(model, "_cfg", None), "cglow_upscale", None)
    H, W = (input_seq.shape[-2] * up, input_seq.shape[-1] * up) if up else (None, None)
    key = init_states(seeds.reshape(-1), [H, W])
    h = key
    out_mu = out_std = None
    if unnormalise and getattr(model, "out_mu", None) is not None:
        out_mu = model.out_mu.to(dev).view(1, -1, 1, 1)
        out_std = model.out_std.to(dev).view(1, -1, 1, 1)
        if float(out_std.abs().sum()) == 0.0:
            out_mu = out_std = None
    frames = []
    for t in range(tmax):
        x = input_seq[:, t].unsqueeze(0).expand(samples, *input_seq[:, t].shape).reshape(samples * B, *input_seq.shape[2:])
        y, _, h = sampler(x, h)
        if t % stride == 0:
            frames.append(y if out_mu is None else out_std * y + out_mu)
        if state_mix_every and t % state_mix_every == 0:
            h = mix_states(h, key)
    yp = torch.stack(frames, 1)                                   # [samples*B, T', C, H, W]
    return yp.reshape(samples, B, *yp.shape[1:])


@torch.no_grad()
def test_error(model, input_seq: torch.Tensor, target_seq: torch.Tensor, samples: int, tmax: int = 40, **kw):
    """The summand of ``TrainFlow.test`` (``nn/trainFlowParallel.py:374``) for one mini-batch: squared error between the mean
    over the samples and the (un-normalised) target over time steps ``1..tmax``, summed.  The reference divides the total by
    ``ntest * tmax * H * W`` (``:377``).  ``target_seq [B, >= tmax + 1, C, H, W]`` already un-normalised."""
    yp = model_pred(model, input_seq, samples, tmax + 1, state_mix_every=kw.pop("state_mix_every", 10), **kw)
    return torch.pow(yp[:, :, 1:tmax + 1].mean(0) - target_seq[:, 1:tmax + 1], 2).sum()


class GraphedSampler:
    """``model.sample`` captured once as CUDA graphs and replayed (reference loop: the 48 flow steps of
    ``LSTMCFlowDecoder.reverse``, nn/tmGlow.py:269-303, called once per time step by every prediction loop).  One replay =
    one ``cudaGraphLaunch`` instead of ~100 kernel launches plus the Python work of a call -- what bounds the latency of a
    small batch.  The ConvLSTM states chain from call to call inside the object (two graphs ping-pong between two state
    buffers, so no state is ever copied); the noise is drawn inside the graph (graph-safe CUDA generator: fresh draws every
    replay, same distribution and order as ``sample``).

    ``x``: the LF input of the first call -- ``[B,nic,h,w]``, or a batch-expanded view ``x1.expand(S, ...)`` for S samples
    of one input (shared-input fast path).  ``sample(x_new)`` copies ``x_new`` (same shape; for the shared case ``[1,nic,h,w]``)
    into the static input and replays; it returns ``(y, log_det, states)`` -- STATIC tensors, overwritten by the next call.
    ``states`` / ``set_states`` give access to the carried LSTM states (e.g. for the reference's periodic state mixing)."""

    def __init__(self, model, x, h_in, warmup=2, eps=None):
        """``eps``: optional explicit noise (list as ``reconstruct`` takes it): the graphs then replay ``model.reconstruct`` on
        static noise buffers that ``sample(eps=...)`` refreshes -- deterministic, bit-comparable with the eager call."""
        assert not model.training, "GraphedSampler is an inference tool (eval mode)"
        self.model = model
        self._eps = [e.detach().clone() for e in eps] if eps is not None else None
        dev = x.device
        shared = x.dim() == 4 and x.shape[0] > 1 and x.stride(0) == 0
        self._S = x.shape[0]
        self._shared = shared
        self._x = (x[:1] if shared else x).detach().clone()
        cl = lambda t: t.detach().clone(memory_format=torch.channels_last) if t.dim() == 4 else t.detach().clone()
        self._h = [[(cl(a), cl(c)) for a, c in h_in], None]
        self._cur = 0
        self.replays = 0
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(warmup, 1)):
                y, ld, h1 = self._call(self._h[0])
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self._h[1] = [(torch.empty_like(a), torch.empty_like(c)) for a, c in h1]
        self._g, self._out = [], []
        pool = None
        for k in (0, 1):                    # graph k reads state buffer k and leaves the new states in buffer 1 - k
            g = torch.cuda.CUDAGraph()
            with torch.no_grad(), (torch.cuda.graph(g) if pool is None else torch.cuda.graph(g, pool=pool)):
                y, ld, hn = self._call(self._h[k])
                for (dh, dc), (sh, sc) in zip(self._h[1 - k], hn):
                    dh.copy_(sh); dc.copy_(sc)
            pool = g.pool()
            self._g.append(g); self._out.append((y, ld))

    def _call(self, h):
        if self._eps is not None:
            return self.model.reconstruct(self._xin(), h, self._eps)
        return self.model.sample(self._xin(), h)

    def _xin(self):
        return self._x.expand(self._S, -1, -1, -1) if self._shared else self._x

    @property
    def states(self):
        return self._h[self._cur]

    def set_states(self, h):
        for (dh, dc), (sh, sc) in zip(self._h[self._cur], h):
            dh.copy_(sh); dc.copy_(sc)

    @torch.no_grad()
    def sample(self, x=None, eps=None):
        if eps is not None:
            for d, e in zip(self._eps, eps):
                d.copy_(e, non_blocking=True)
        if x is not None:
            self._x.copy_(x[:1] if (self._shared and x.shape[0] != 1) else x, non_blocking=True)
        k = self._cur
        self._g[k].replay()
        self._cur = 1 - k
        self.replays += 1
        y, ld = self._out[k]
        return y, ld, self._h[self._cur]
