"""The reference's multi-GPU caller pattern: ``DataParallelINNModel`` replicates the module with
``nn.DataParallel.replicate`` and calls ``replica.sample(*input)`` from ONE PYTHON THREAD PER GPU inside
``torch.cuda.device(dev)`` (utils/parallel.py:150-169, 174-241).  The drop-in must keep working under it: every replica
owns its library handle, flat parameter copy and workspace, and the library is re-entrant.

Runs on two GPUs when the box has them, otherwise with two replicas (two threads) on the same device."""
import json
import threading

import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _parallel_apply(replicas, inputs, devices, method):
    """The thread pattern of inn_parallel_apply (utils/parallel.py:204-241)."""
    lock, results = threading.Lock(), {}

    def worker(i, module, inp, device):
        try:
            with torch.no_grad(), torch.cuda.device(device):
                out = getattr(module, method)(*inp)
            with lock:
                results[i] = out
        except Exception as ex:          # re-raised in the caller, like ExceptionWrapper.reraise()
            with lock:
                results[i] = ex
    threads = [threading.Thread(target=worker, args=(i, m, inp, d)) for i, (m, inp, d) in enumerate(zip(replicas, inputs, devices))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for i in range(len(replicas)):
        if isinstance(results[i], Exception):
            raise results[i]
    return [results[i] for i in range(len(replicas))]


@pytest.mark.parametrize("precision", ["fp32", "f16x3"])
def test_replicate_and_one_thread_per_device(precision):
    from oracle import tmglow_oracle as O
    from tmglow_b200 import TMGlow
    g = load_golden("caseA_states")
    cfg = json.loads(g["config"])
    m = TMGlow(cfg["in_features"], cfg["out_features"], cfg["enc_blocks"], cfg["glow_blocks"],
               cond_features=cfg["cond_features"], cglow_upscale=cfg["cglow_upscale"], growth_rate=cfg["growth_rate"],
               init_features=cfg["init_features"], rec_features=cfg["rec_features"])
    m.load_state_dict(g["state_dict"])
    m = m.to("cuda:0").eval()
    m.precision = precision
    devices = [0, 1] if torch.cuda.device_count() >= 2 else [0, 0]
    with torch.no_grad():
        replicas = torch.nn.parallel.replicate(m, devices, detach=True)
    assert all(r is not m and r._handles is not m._handles for r in replicas)
    ocfg = O.OracleConfig.from_dict(cfg)
    x, eps, h_in = g["x"], g["rec2"]["eps"], g["h_in"]
    B = x.shape[0]
    assert B >= 2
    y_o, ld_o, h_o = O.reconstruct(g["state_dict"], ocfg, x, h_in, eps)
    # scatter along the batch dimension, one chunk per replica (DataParallel.scatter)
    cuts = [(0, B // 2), (B // 2, B)]
    inputs = []
    for (lo, hi), d in zip(cuts, devices):
        dev = torch.device("cuda", d)
        inputs.append((x[lo:hi].to(dev), [(a[lo:hi].to(dev), c[lo:hi].to(dev)) for a, c in h_in], [e[lo:hi].to(dev) for e in eps]))
    for rounds in range(3):                                   # repeated concurrent calls: caches are per replica
        outs = _parallel_apply(replicas, inputs, devices, "reconstruct")
        for (lo, hi), (y, ld, h), d in zip(cuts, outs, devices):
            assert y.device.index == d
            assert (y.cpu() - y_o[lo:hi]).abs().max().item() < 2e-4
            assert ((ld.cpu() - ld_o[lo:hi]).abs() / ld_o[lo:hi].abs()).max().item() < 1e-5
            for l, (a, c) in enumerate(h):
                assert (a.cpu() - h_o[l][0][lo:hi]).abs().max().item() < 2e-4
    # the original still works and still owns its parameters after the replicas ran
    dev0 = torch.device("cuda:0")
    y, ld, _ = m.reconstruct(x.to(dev0), [(a.to(dev0), c.to(dev0)) for a, c in h_in], [e.to(dev0) for e in eps])
    assert (y.cpu() - y_o).abs().max().item() < 2e-4
    # sample() through the same thread pattern (the call the reference makes, parallel.py:213): shapes and devices
    outs = _parallel_apply(replicas, [(i[0], i[1]) for i in inputs], devices, "sample")
    for (lo, hi), (y, ld, h) in zip(cuts, outs):
        assert tuple(y.shape) == (hi - lo,) + tuple(y_o.shape[1:]) and torch.isfinite(y).all()


@pytest.mark.parametrize("shared", [False, True])
def test_graphed_sampler_replays_equal_eager_calls(shared):
    """uq.GraphedSampler: the whole sample()/reconstruct() call captured as CUDA graphs (two graphs ping-pong between two LSTM
    state buffers).  With explicit noise the replays must reproduce the eager calls bit for bit over several chained time
    steps (same kernels, same order); with internal noise the states chain and every replay draws fresh noise."""
    from tmglow_b200 import TMGlow, uq
    g = load_golden("caseA_states")
    cfg = json.loads(g["config"])
    m = TMGlow(cfg["in_features"], cfg["out_features"], cfg["enc_blocks"], cfg["glow_blocks"],
               cond_features=cfg["cond_features"], cglow_upscale=cfg["cglow_upscale"], growth_rate=cfg["growth_rate"],
               init_features=cfg["init_features"], rec_features=cfg["rec_features"])
    m.load_state_dict(g["state_dict"])
    dev = torch.device("cuda:0")
    m = m.to(dev).eval()
    m.precision = "f16x3"
    B = g["x"].shape[0]
    x = g["x"].to(dev)
    xs = [x[:1].expand(B, -1, -1, -1) if shared else x, (1.3 * x[:1]).expand(B, -1, -1, -1) if shared else 1.3 * x, 0.7 * x[:1] if shared else 0.7 * x]
    if shared:
        xs[2] = xs[2].expand(B, -1, -1, -1)
    h0 = [(a.to(dev), c.to(dev)) for a, c in g["h_in"]]
    eps = [[(e * (1.0 + 0.2 * t)).to(dev) for e in g["rec2"]["eps"]] for t in range(3)]
    ref, h = [], h0
    for t in range(3):
        y, ld, h = m.reconstruct(xs[t], h, eps[t])
        ref.append((y.clone(), ld.clone()))
    gs = uq.GraphedSampler(m, xs[0], h0, eps=eps[0])
    gs.set_states(h0)                                   # the warm-up calls advanced nothing: states restart from h0
    for t in range(3):
        y, ld, hs = gs.sample(xs[t][:1] if shared else xs[t], eps=eps[t])
        assert torch.equal(y, ref[t][0]) and torch.equal(ld, ref[t][1])
    for (a, c), (ar, cr) in zip(hs, h):
        assert torch.equal(a, ar) and torch.equal(c, cr)
    assert gs.replays == 3
    gn = uq.GraphedSampler(m, xs[0], h0)                # internal noise
    y1 = gn.sample()[0].clone()
    y2 = gn.sample()[0].clone()
    assert torch.isfinite(y1).all() and torch.isfinite(y2).all() and not torch.equal(y1, y2)
