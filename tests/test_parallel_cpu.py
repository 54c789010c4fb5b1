"""Host-side logic of the N > 1 (sample-parallel) path on CPU: world_size-2 gloo processes.

The CUDA path cannot run here; what shards samples over ranks, seeds the LSTM states, streams the per-time-step
moments and combines them with ONE all-reduce per sequence (tmglow_b200/uq.py) is plain host code and is exercised
with a CPU stand-in for ``TMGlow.sample`` whose output depends on the per-sample seed, so any sharding or
reduction mistake changes the result.
"""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "deep-turbulence_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _init_states(seeds, dims):
    hs = []
    for sd in seeds.tolist():
        g = torch.Generator().manual_seed(int(sd))
        hs.append(torch.rand(1, 2, 2, 2, generator=g))
    h = torch.cat(hs, 0)
    return [(h, h.clone())]


def _sampler(x, h):
    """y depends on the LF input, on the per-sample state and on a state-dependent 'noise'."""
    (hh, cc), = h
    base = x.mean(dim=(1, 2, 3), keepdim=True)
    y = base + hh.mean(dim=(1, 2, 3)).view(-1, 1, 1, 1) * torch.ones(x.shape[0], 3, 4, 4) + 0.1 * cc.sum(dim=(1, 2, 3)).view(-1, 1, 1, 1)
    h2 = [(0.9 * hh + 0.05, cc + 0.01 * hh)]
    return y, torch.zeros(x.shape[0]), h2


def _run_rank(rank, world, port, samples, out):
    import torch.distributed as dist
    from tmglow_b200 import uq
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    x_seq = torch.randn(5, 4, 2, 2, generator=g)
    mean, var, ntot, kept = uq.sample_sequence(None, x_seq, samples, base_seed=7, rank=rank, world=world,
                                               state_mix_every=2, unnormalise=False, keep_samples=True,
                                               sampler=_sampler, init_states=_init_states)
    if rank == 0:
        torch.save({"mean": mean, "var": var, "n": ntot, "kept": kept.shape[0]}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions():
    from tmglow_b200 import uq
    for total in (0, 1, 7, 8, 1000, 1023):
        for world in (1, 2, 3, 8):
            parts = [uq.shard_range(total, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        uq.shard_range(4, 2, 2)


def test_seeds_depend_on_global_index_only():
    from tmglow_b200 import uq
    full = uq.sample_seeds(11, 3, 0, 10)
    assert torch.equal(torch.cat([uq.sample_seeds(11, 3, 0, 4), uq.sample_seeds(11, 3, 4, 10)]), full)
    assert len(set(full.tolist())) == 10
    assert not torch.equal(uq.sample_seeds(11, 4, 0, 10), full)


@pytest.mark.parametrize("samples", [8, 7])
def test_two_rank_gloo_matches_single_process(tmp_path, samples):
    from tmglow_b200 import uq
    out = str(tmp_path / "r0.pt")
    mp.spawn(_run_rank, args=(2, _free_port(), samples, out), nprocs=2, join=True)
    got = torch.load(out)
    g = torch.Generator().manual_seed(0)
    x_seq = torch.randn(5, 4, 2, 2, generator=g)
    mean, var, ntot, kept = uq.sample_sequence(None, x_seq, samples, base_seed=7, state_mix_every=2, unnormalise=False,
                                               keep_samples=True, sampler=_sampler, init_states=_init_states)
    assert got["n"] == samples == ntot
    assert got["kept"] == (samples + 1) // 2            # rank 0 holds the larger shard
    assert torch.allclose(got["mean"], mean, rtol=0, atol=1e-6)
    assert torch.allclose(got["var"], var, rtol=1e-5, atol=1e-7)
    # the moments are those of the kept samples
    assert torch.allclose(mean, kept.mean(0), atol=1e-5)
    assert torch.allclose(var, kept.var(0, unbiased=True), atol=1e-5)


def test_bench_sharding_is_weak_scaling():
    """bench.py gives every rank its own LF input and S samples (weak scaling, no data-path collective): the
    whole-job value is world * S * K / max-over-ranks time."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    c = bench.workload_config(1024, 8)
    assert c["samples_per_gpu_per_step"] == 1024 and "x8" in c["parallelism"] and "no collective" in c["parallelism"]


def _run_dp_rank(rank, world, port, out):
    import torch.distributed as dist
    from tmglow_b200 import train as T
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    g = torch.Generator().manual_seed(100 + rank)
    grad = torch.randn(1000, generator=g)          # this rank's flat gradient
    T.allreduce_mean_(grad)
    norm = T.clip_flat_grad_(grad, 1.0)
    if rank == 0:
        torch.save({"grad": grad, "norm": norm}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_dp_gradient_allreduce_and_clip_gloo(tmp_path):
    """Data-parallel step logic: mean of the ranks' flat gradients (one all-reduce), then global-norm clipping --
    identical on every rank, equal to the single-process computation on the concatenated batch of equal shards."""
    from tmglow_b200 import train as T
    out = str(tmp_path / "dp.pt")
    mp.spawn(_run_dp_rank, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    gs = [torch.randn(1000, generator=torch.Generator().manual_seed(100 + r)) for r in range(2)]
    ref = (gs[0] + gs[1]) / 2
    n = float(ref.norm())
    assert abs(got["norm"] - n) < 1e-5
    assert torch.allclose(got["grad"], ref * (1.0 / (n + 1e-6)), atol=1e-6)
    assert abs(float(got["grad"].norm()) - 1.0) < 1e-4
    # single process: no process group -> no-op
    t = torch.ones(4)
    assert torch.equal(T.allreduce_mean_(t.clone()), t)


class _ToyModel:
    """CPU stand-in with the host-side surface train_block / train_series use (the CUDA model has no CPU path): a flat
    parameter buffer whose first 4 entries are trainable and last 2 are 'buffers', sample_train = a differentiable toy."""
    glow_blocks = [1]

    def __init__(self):
        self._flat = torch.nn.Parameter(torch.tensor([0.5, -0.25, 1.0, 2.0, 7.0, 9.0]))
        self.flat_grad = torch.zeros(6)
        self.finalized = 0
        self.refreshed = 0

    def flat_parameter_for_optimizer(self):
        return self._flat

    def trainable_mask(self):
        return torch.tensor([1.0, 1.0, 1.0, 1.0, 0.0, 0.0])

    def zero_flat_grad(self):
        self.flat_grad.zero_()

    def finalize_flat_grad(self):
        self.finalized += 1

    def refresh_weights(self):
        self.refreshed += 1

    def initLSTMStates(self, seeds, dims):
        return [(torch.ones(len(seeds), 1, 1, 1), 2 * torch.ones(len(seeds), 1, 1, 1))]

    def sample_train(self, x, h):
        w = self._flat[:4].detach()
        hh, cc = h[0]
        y = x * w[0] + hh * w[1]
        self.flat_grad[:4] += torch.tensor([1.0, 2.0, 3.0, 4.0])       # what the CUDA backward would accumulate
        return y.requires_grad_(True), y.flatten(1).sum(1), hh + 1.0, cc * 0.5


def test_train_block_host_logic_weight_decay_and_finalize():
    """train_block: finalize before the all-reduce, clip, then the reference's Adam weight decay (main.py:78) added to the
    TRAINABLE entries only (buffers inside the flat parameter buffer keep a zero gradient), optimizer step, refresh."""
    from tmglow_b200 import train as T
    m = _ToyModel()
    opt = torch.optim.SGD([m.flat_parameter_for_optimizer()], lr=0.1)
    x = torch.ones(2, 3, 1, 1, 1)
    before = m._flat.detach().clone()
    loss, norm, h = T.train_block(m, opt, x, torch.zeros(2, 3, 1, 1, 1), m.initLSTMStates(torch.arange(2), None),
                                  loss_fn=lambda y, ld, t: (y ** 2).mean() + 0 * ld.sum(), max_norm=1.0, weight_decay=0.5)
    assert m.finalized == 1 and m.refreshed == 1
    g = torch.tensor([3.0, 6.0, 9.0, 12.0, 0.0, 0.0])                   # three time steps of the toy gradient
    assert abs(norm - float(g.norm())) < 1e-5
    expect = g / (float(g.norm()) + 1e-6) + 0.5 * before * m.trainable_mask()
    assert torch.allclose(m._flat.detach(), before - 0.1 * expect, atol=1e-6)
    assert torch.equal(m._flat.detach()[4:], before[4:])                # buffers untouched
    assert not h[0][0].requires_grad


def test_train_series_blocks_and_state_mixing():
    """train_series = one mini-batch of TrainFlow.trainParallel (trainFlowParallel.py:225-303): Tmax // tback optimizer
    steps, states mixed 50/50 with the initial states after each block, losses summed."""
    from tmglow_b200 import train as T
    m = _ToyModel()
    opt = torch.optim.SGD([m.flat_parameter_for_optimizer()], lr=0.0)
    calls = []

    def crit(y, lp, tgt, tm, tr):
        calls.append((tuple(y.shape), tuple(tm.shape), tuple(tr.shape)))
        return (y ** 2).mean()
    x0 = torch.ones(2, 6, 1, 1, 1)
    tot, a = T.train_series(m, opt, crit, x0, torch.zeros(2, 6, 1, 1, 1), torch.arange(2), tback=3, max_norm=None)
    assert len(calls) == 2 and calls[0] == ((2, 3, 1, 1, 1), (2, 1, 1, 1), (2, 1, 1, 1))
    assert m.refreshed == 2
    # h: 1 -> +3 -> mix with 1 -> (4+1)/2 = 2.5 -> +3 -> (5.5+1)/2 = 3.25 ; c: 2 -> /8 -> (0.25+2)/2 = 1.125 -> /8 -> mix
    assert torch.allclose(a[0][0], torch.full((2, 1, 1, 1), 3.25))
    assert torch.allclose(a[0][1], torch.full((2, 1, 1, 1), (1.125 / 8 + 2) / 2))
    assert tot.ndim == 0 and float(tot) > 0
    ms = T.mix_states([(torch.ones(1), torch.zeros(1))], [(3 * torch.ones(1), 2 * torch.ones(1))])
    assert float(ms[0][0]) == 2.0 and float(ms[0][1]) == 1.0


def test_model_pred_matches_reference_loops():
    """uq.model_pred folds the reference's sample loop (utils.py:197-222 / trainFlowParallel.py:345-367) into the batch:
    same predictions as the loop over samples with the same seeds, state mixing at the same time steps."""
    from tmglow_b200 import uq
    g = torch.Generator().manual_seed(2)
    Bc, T, S = 3, 7, 4
    inp = torch.randn(Bc, T, 4, 2, 2, generator=g)
    seeds = torch.arange(S * Bc).reshape(S, Bc) + 50
    model = type("M", (), {"out_mu": None})()
    got = uq.model_pred(model, inp, S, T, stride=2, state_mix_every=3, seeds=seeds, sampler=_sampler, init_states=_init_states)
    assert got.shape[:3] == (S, Bc, (T + 1) // 2)
    for i in range(S):                                       # the reference's loops, literally
        hkey = _init_states(seeds[i], None)
        h0 = hkey
        k = 0
        for t in range(T):
            y, _, h0 = _sampler(inp[:, t], h0)
            if t % 2 == 0:
                assert torch.allclose(got[i, :, k], y, atol=1e-6), (i, t)
                k += 1
            if t % 3 == 0:
                h0 = [(0.5 * a + 0.5 * ak, 0.5 * c + 0.5 * ck) for (a, c), (ak, ck) in zip(h0, hkey)]
    tgt = torch.randn(Bc, T, *got.shape[3:], generator=g)
    err = uq.test_error(model, inp, tgt, S, tmax=T - 1, seeds=seeds, sampler=_sampler, init_states=_init_states)
    full = uq.model_pred(model, inp, S, T, state_mix_every=10, seeds=seeds, sampler=_sampler, init_states=_init_states)
    assert torch.allclose(err, ((full[:, :, 1:].mean(0) - tgt[:, 1:]) ** 2).sum())
