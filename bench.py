#!/usr/bin/env python
"""bench.py -- TM-Glow hot-path benchmark (driver contract in the task statement, tier section 4).

Workload (BASELINE.json configs[1]): backward-facing-step inference, default model
(enc [4,4,4], glow [16,16,16], 1 746 573 parameters, random init + the well-conditioned
perturbation of SURVEY.md appendix B), ONE low-fidelity snapshot x[1,4,32,64] per GPU and S
stochastic high-fidelity samples y[S,3,64,128] per step through ``TMGlow.sample`` with the ConvLSTM
states carried from step to step.  One "step" = one ``sample()`` call over the S samples.
metric = HF samples/s, whole job (all ranks).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--samples S] [--impl b200|reference]

N > 1 is launched by the driver with torch.distributed.run (one rank per GPU); samples are sharded
over ranks with no data-path collective (weak scaling: S per GPU), NCCL is used only for the
barrier and the max-over-ranks of the device-timed duration.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "deep-turbulence_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

GEOM = dict(nic=4, h=32, w=64, noc=3, H=64, W=128)
MODEL_KW = dict(cond_features=32, cglow_upscale=2, growth_rate=4, init_features=16, rec_features=64)
# SURVEY.md 8(d) / BASELINE.md section 3: algorithmic work per HF sample, backward-step geometry
ALG_FLOP_PER_SAMPLE = 2.159e9
ALG_BYTES_PER_SAMPLE = 13.9e6
DTYPES = {"fp32": "f32", "tf32x3": "f32 (3xTF32 tensor-core split)", "tf32": "tf32",
          "f16x3": "f32 (fp16 hi+lo tensor-core split, fp32 accumulate)", "f16": "f16 operands, f32 accumulate"}


def perturb_(model, seed):
    """SURVEY.md appendix B recipe (zc = 0.002): non-trivial couplings/priors, well conditioned."""
    gen = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            r = torch.randn(p.shape, generator=gen)
            if name.endswith("norm.weight"):
                p.copy_(torch.exp(0.1 * r))
            elif name.endswith("norm.bias"):
                p.copy_(0.1 * r)
            elif name.endswith("conv.log_s"):
                p.add_(0.05 * r)
            elif name.endswith(".scale"):
                p.copy_(0.1 * r)
            elif "zero_conv.conv." in name or "latent_encoder.conv2d.conv." in name:
                p.copy_(0.002 * r)


def build_model():
    from tmglow_b200 import TMGlow
    torch.manual_seed(12345)       # args.py:151
    np.random.seed(12345)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        m = TMGlow(GEOM["nic"], GEOM["noc"], [4, 4, 4], [16, 16, 16], **MODEL_KW)
    perturb_(m, 12346)
    m.eval()
    return m


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "20", "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return None
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "samples": len(sm), "reasons": sorted(reasons)}


class stdout_to_stderr:
    """File-descriptor level redirect: stdout must carry exactly one JSON line, native libraries (NCCL) write there too."""
    def __enter__(self):
        sys.stdout.flush()
        self._saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self._saved, 1)
        os.close(self._saved)
        return False


def reference_model(geom, kw, train=False):
    """The UNMODIFIED reference ``TMGlow`` (oracle/_ref, made by oracle/make_ref.py) on the CPU with the reference's own
    initialisation under the bench seeds plus the same well-conditioned perturbation as the GPU arm.  Returns
    ``(model, "reference")`` or ``(None, why)`` when no copy of the reference is available.  Nothing of tmglow_b200 is
    imported on this path."""
    try:
        from oracle import ref_loader
        ns = ref_loader.load(trainer=train)
        if train:
            ref_loader.grad_shim()
    except ImportError as ex:
        return None, str(ex)
    import contextlib
    import io
    torch.manual_seed(12345); np.random.seed(12345)
    with contextlib.redirect_stdout(io.StringIO()):
        m = ns.TMGlow(geom["nic"], geom["noc"], [4, 4, 4], [16, 16, 16], **kw)
    perturb_(m, 12346)
    m.train() if train else m.eval()
    m._ref_ns = ns
    return m, "reference"


class _CpuSampler:
    """sample() of the path on the host cores: the real reference when oracle/_ref exists (kind "reference"), else the
    pinned oracle port (kind "port"; it runs the same ATen CPU operators in the same order)."""

    def __init__(self):
        self.threads = os.cpu_count() or 1
        torch.set_num_threads(self.threads)
        self.ref, why = reference_model(GEOM, MODEL_KW)
        self.kind = "reference" if self.ref is not None else "port"
        if self.ref is None:
            from oracle import tmglow_oracle as O
            m = build_model()            # only for its random-init state_dict; never moved to a GPU, no kernel runs
            self.O, self.sd, self.cfg = O, {k: v.detach().clone() for k, v in m.state_dict().items()}, O.OracleConfig.from_dict(m._cfg_dict)
            self.why = why
        self.gen = torch.Generator().manual_seed(1)

    def states(self, batch):
        if self.ref is not None:
            return self.ref.initLSTMStates(torch.arange(batch), [GEOM["H"], GEOM["W"]])
        return self.O.init_lstm_states(self.cfg, torch.arange(batch), [GEOM["H"], GEOM["W"]])

    def x(self, batch):
        return torch.randn(1, GEOM["nic"], GEOM["h"], GEOM["w"], generator=self.gen).expand(batch, -1, -1, -1).contiguous()

    def sample(self, x, h):
        with torch.no_grad():
            if self.ref is not None:
                return self.ref.sample(x, h)
            return self.O.sample(self.sd, self.cfg, x, h, self.gen)

    def rate(self, batch, calls, warm=1):
        x, h = self.x(batch), self.states(batch)
        for _ in range(warm):
            y, ld, h = self.sample(x, h)
        t0 = time.perf_counter()
        for _ in range(calls):
            y, ld, h = self.sample(x, h)
        return batch * calls / (time.perf_counter() - t0)

    def best_batch(self, candidates=(16, 64, 256)):
        """samples/s grows with the batch until the cores saturate: a short sweep, stated in the report."""
        sweep = {}
        for b in candidates:
            sweep[b] = self.rate(b, 1)
        best = max(sweep, key=sweep.get)
        return best, sweep


def cpu_samples_per_s(min_seconds, max_calls=400):
    """cpu_baseline leg: the path on the host cores at the batch where samples/s saturates, for about `min_seconds`."""
    s = _CpuSampler()
    batch, sweep = s.best_batch()
    x, h = s.x(batch), s.states(batch)
    y, ld, h = s.sample(x, h)
    n, t0 = 0, time.perf_counter()
    while True:
        y, ld, h = s.sample(x, h)
        n += 1
        el = time.perf_counter() - t0
        if el >= min_seconds or n >= max_calls:
            break
    return batch * n / el, s, batch, n, el, sweep


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on all host cores; rank 0 only.  The REAL
    ``TMGlow.sample`` (oracle/_ref: an unmodified copy of the reference package, kind "reference"); only when that copy is
    missing, the pinned oracle port (kind "port").  Each step is a bounded sample of the GPU arm's step (S = 4096 samples of
    one LF input): `batch` samples, with `batch` taken where samples/s saturates on this host (sweep in the line)."""
    if rank != 0:
        return
    s = _CpuSampler()
    if args.ref_batch > 0:
        B, sweep = args.ref_batch, None
    else:
        B, sweep = s.best_batch()
    x, h = s.x(B), s.states(B)
    for _ in range(args.warmup):
        y, ld, h = s.sample(x, h)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        y, ld, h = s.sample(x, h)
    el = time.perf_counter() - t0
    val = B * args.steps / el
    line = {
        "impl": "reference", "metric": "hf_samples_per_sec", "value": val, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(workload_config(args.samples, args.gpus), precision="fp32 (reference, CPU)"),
        "cpu_baseline": {"value": val, "unit": "samples/s", "cores": s.threads, "kind": s.kind,
                         "sample": "%d steps x %d of the %d HF samples of a step through %s on %d host threads%s" % (
                             args.steps, B, args.samples,
                             "the unmodified reference TMGlow.sample (oracle/_ref)" if s.kind == "reference" else "oracle.sample()",
                             s.threads, "; samples/s by batch: %s" % {k: round(v, 1) for k, v in sweep.items()} if sweep else "")},
        "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(S, n, note=None):
    c = {"workload": "TM-Glow backward-facing-step inference: 1 LF input x %d stochastic HF samples per GPU per step "
                     "(BASELINE.json configs[1]); x[1,4,32,64] -> y[S,3,64,128]; default model 1746573 params" % S,
         "samples_per_gpu_per_step": S, "parallelism": "sample-parallel x%d, no collective" % n,
         "l2_policy": "working set per step (%.1f GB of activations) far exceeds the 126 MB L2; no explicit flush" %
                      (S * 3.5e6 / 1e9)}
    if note:
        c["note"] = note
    return c


TRAIN_GEOM = dict(nic=3, h=16, w=16, noc=3, H=64, W=64)       # cylinder-array (x[B,3,16,16] -> y[B,3,64,64])
TRAIN_KW = dict(cond_features=32, cglow_upscale=4, growth_rate=4, init_features=16, rec_features=64)


def cpu_port_train_steps(batch, tback, steps, warmup=0):
    """One optimizer step of the reference training objective on the host cores: BPTT block of `tback` sample() calls under
    autograd (training-mode BatchNorm), TMGLowLoss, clip, Adam-amsgrad -- the arithmetic of trainFlowParallel.py:248-300 at
    a bounded batch.  Runs the UNMODIFIED reference classes (oracle/_ref: ``TMGlow.sample``, ``TMGLowLoss``; the only
    accommodation is the out-of-place clamp of SURVEY 8c so that autograd accepts ``GaussianDiag``) -- kind "reference";
    only when that copy is missing, the pinned oracle ports (kind "port").
    Returns (seconds per step, threads, loss, kind)."""
    import types
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    ref, _ = reference_model(TRAIN_GEOM, TRAIN_KW, train=True)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(batch, tback, TRAIN_GEOM["nic"], TRAIN_GEOM["h"], TRAIN_GEOM["w"], generator=g)
    tgt = torch.randn(batch, tback, TRAIN_GEOM["noc"], TRAIN_GEOM["H"], TRAIN_GEOM["W"], generator=g)
    t_mean = tgt.mean(1)
    t_rms = torch.sqrt(torch.mean((tgt - t_mean.unsqueeze(1)) ** 2, dim=1))          # trainFlowParallel.py:237-238
    if ref is not None:
        ref.out_mu, ref.out_std = torch.zeros(3), torch.ones(3)
        crit = ref._ref_ns.TMGLowLoss(types.SimpleNamespace(beta=200.0, dx=5.0 / 64, dy=5.0 / 64), types.SimpleNamespace(module=ref))
        params = list(ref.parameters())
        opt = torch.optim.Adam(params, lr=1e-4, weight_decay=1e-8, amsgrad=True)       # main.py:78
        h_key = ref.initLSTMStates(torch.arange(batch), [TRAIN_GEOM["H"], TRAIN_GEOM["W"]])
        h = h_key
        el, loss = 0.0, None
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            ys, lds = [], []
            for t in range(tback):
                y, ld, h = ref.sample(x[:, t], h)
                ys.append(y); lds.append(ld)
            loss = crit(torch.stack(ys, 1), torch.stack(lds, 1), tgt, t_mean, t_rms)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(params, 1.0)
            opt.step()
            opt.zero_grad()
            h = [(0.5 * a.detach() + 0.5 * ak, 0.5 * c.detach() + 0.5 * ck) for (a, c), (ak, ck) in zip(h, h_key)]
            if it >= warmup:
                el += time.perf_counter() - t0
        return el / steps, threads, float(loss.detach()), "reference"
    from oracle import tmglow_oracle as O
    from oracle import tmglow_loss_oracle as OL
    from tmglow_b200 import TMGlow
    torch.manual_seed(12345); np.random.seed(12345)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        m = TMGlow(TRAIN_GEOM["nic"], TRAIN_GEOM["noc"], [4, 4, 4], [16, 16, 16], **TRAIN_KW)
    perturb_(m, 12346)
    cfg = O.OracleConfig.from_dict(m._cfg_dict)
    trainable = {n for n, _ in m.named_parameters()}
    sd = {k: (v.detach().clone().requires_grad_(True) if k in trainable else v.detach().clone()) for k, v in m.state_dict().items()}
    params = [v for k, v in sd.items() if k in trainable]
    opt = torch.optim.Adam(params, lr=1e-4, weight_decay=1e-8, amsgrad=True)
    mu, sdv = torch.zeros(3), torch.ones(3)
    h_key = O.init_lstm_states(cfg, torch.arange(batch), [TRAIN_GEOM["H"], TRAIN_GEOM["W"]])
    h = h_key
    el, loss = 0.0, None
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad()
        ys, lds = [], []
        for t in range(tback):
            eps = O.draw_eps(cfg, batch, TRAIN_GEOM["H"], TRAIN_GEOM["W"], g)
            y, ld, h = O.reconstruct(sd, cfg, x[:, t], h, eps, training=True)
            ys.append(y); lds.append(ld)
        loss = OL.tmglow_loss(torch.stack(ys, 1), torch.stack(lds, 1), tgt, t_rms, mu, sdv, 5.0 / 64, 5.0 / 64, 200.0)
        loss.backward()
        torch.nn.utils.clip_grad_norm_([p_ for p_ in params if p_.grad is not None], 1.0)
        opt.step()
        h = [(0.5 * a.detach() + 0.5 * ak, 0.5 * c.detach() + 0.5 * ck) for (a, c), (ak, ck) in zip(h, h_key)]
        if it >= warmup:
            el += time.perf_counter() - t0
    return el / steps, threads, float(loss.detach()), "port"


def run_reference_train(args, rank, world):
    """--impl reference --workload train: the reference's training step on the host cores (oracle port), bounded batch."""
    if rank != 0:
        return
    Bc = args.ref_train_batch
    sec, threads, loss, kind = cpu_port_train_steps(Bc, args.tback, max(args.steps, 1), warmup=min(args.warmup, 1))
    val = 1.0 / sec
    line = {"impl": "reference", "metric": "train_steps_per_sec", "value": val, "unit": "steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "TM-Glow cylinder-array training (BASELINE.json configs[2]) on the CPU: bounded sample, batch %d "
                                   "(the GPU arm's step is global batch %d), BPTT block of %d time steps, reference TMGLowLoss, "
                                   "clip 1.0, Adam-amsgrad" % (Bc, args.global_batch, args.tback),
                       "global_batch": Bc, "tback": args.tback, "parallelism": "cpu"},
            "cpu_baseline": {"value": val, "unit": "steps/s", "cores": threads, "kind": kind,
                             "sample": "%d optimizer step(s) at batch %d x %d time steps through %s (autograd)" % (
                                 max(args.steps, 1), Bc, args.tback,
                                 "the unmodified reference TMGlow.sample + TMGLowLoss (oracle/_ref)" if kind == "reference" else "the oracle port")},
            "e2e": {"value": val, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "hf_snapshots_per_sec": Bc * args.tback / sec, "gpu_launches": 0, "loss": loss}
    print(json.dumps(line))


def measure_train(args, rank, world, dev, dist, steps, warmup, e2e_steps=0):
    """Times `steps` optimizer steps (after `warmup`) of data-parallel training with device-resident inputs and, when
    e2e_steps > 0, the same step fed from pinned host memory; returns a dict (ms, launches, loss, norm, e2e_*)."""
    from tmglow_b200 import TMGlow, _lib, train as T
    lib = _lib.load()
    torch.manual_seed(12345); np.random.seed(12345)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        m = TMGlow(TRAIN_GEOM["nic"], TRAIN_GEOM["noc"], [4, 4, 4], [16, 16, 16], **TRAIN_KW)
    perturb_(m, 12346)
    m = m.to(dev).train()
    m.precision = args.precision
    GB, tb = args.global_batch, args.tback
    assert GB % world == 0
    Bl = GB // world
    g = torch.Generator().manual_seed(7 + rank)
    x = torch.randn(Bl, tb, TRAIN_GEOM["nic"], TRAIN_GEOM["h"], TRAIN_GEOM["w"], generator=g).to(dev)
    tgt = torch.randn(Bl, tb, TRAIN_GEOM["noc"], TRAIN_GEOM["H"], TRAIN_GEOM["W"], generator=g).to(dev)
    h = m.initLSTMStates(torch.arange(Bl) + 1000 * rank, [TRAIN_GEOM["H"], TRAIN_GEOM["W"]])
    h_key = h
    # main.py:78 (Adam, amsgrad, weight_decay 1e-8 -- applied by train_block on the trainable entries only); lr below the
    # reference's 1e-3 start because the synthetic targets are white noise -- the arithmetic per step is identical
    from tmglow_b200 import FlatAdam
    opt = FlatAdam(m, lr=1e-4, amsgrad=True)          # clip + decay + AMSGrad fused, no host synchronisation per step
    # the reference loss: TMGLowLoss(beta=200, dx=dy=5/64) with the PDE-residual terms (args.py:61-63), one fused kernel
    import types
    from tmglow_b200.loss import TMGLowLoss, target_statistics
    crit = TMGLowLoss(types.SimpleNamespace(beta=200.0, dx=5.0 / 64, dy=5.0 / 64), m).to(dev)
    t_mean, t_rms = target_statistics(tgt)

    def block(h, xb, tb_, tm, tr):
        loss, norm, h_out = T.train_block(m, opt, xb, tb_, h, max_norm=1.0, criterion=crit, target_mean=tm, target_rms=tr,
                                          weight_decay=1e-8)
        return loss, norm, T.mix_states(h_out, h_key)          # trainFlowParallel.py:296-300

    def timed(fn, n):
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        lib.tmg_launch_count(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        launches = lib.tmg_launch_count(0)
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, int(launches)

    state = {"h": h, "loss": None, "norm": None}
    graphed = None
    if not getattr(args, "no_graph", False):
        try:      # the whole optimizer step as CUDA graphs (tmglow_b200.train.GraphedTrainBlock); eager when capture fails
            graphed = T.GraphedTrainBlock(m, opt, crit, x, tgt, t_mean, t_rms, h_key, max_norm=1.0, weight_decay=1e-8)
        except Exception as ex:
            sys.stderr.write("CUDA-graph capture of the training step failed (%r): running eagerly\n" % (ex,))
            graphed = None

    def resident():
        if graphed is not None:
            state["loss"], state["norm"] = graphed.step()
        else:
            state["loss"], state["norm"], state["h"] = block(state["h"], x, tgt, t_mean, t_rms)

    for _ in range(max(warmup, 1)):
        resident()
    if graphed is not None:
        graphed.time_allreduce = True
        graphed.allreduce_ms()
    ms, launches = timed(resident, steps)
    out = {"ms": ms, "launches": launches + (graphed.kernels_per_step * steps if graphed is not None else 0)}
    if graphed is not None:
        ar_ms, ar_n = graphed.allreduce_ms()
        graphed.time_allreduce = False
        out["allreduce_ms_per_step"] = ar_ms / max(ar_n, 1) if ar_n else 0.0
    if e2e_steps > 0:
        # end to end: every step copies its inputs (LF block + HF targets) from pinned host memory, recomputes the target
        # statistics and reads the loss back
        xh, th = x.cpu().pin_memory(), tgt.cpu().pin_memory()

        def e2e():
            if graphed is not None:
                graphed.load(x_block=xh, target=th)                      # pinned host -> static device buffers
                tm, tr = target_statistics(graphed.target)
                graphed.load(target_mean=tm, target_rms=tr)
                state["loss"], state["norm"] = graphed.step()
            else:
                xb, tb_ = xh.to(dev, non_blocking=True), th.to(dev, non_blocking=True)
                tm, tr = target_statistics(tb_)
                state["loss"], state["norm"], state["h"] = block(state["h"], xb, tb_, tm, tr)
            state["loss_host"] = float(state["loss"])
        e2e()
        ms_e, _ = timed(e2e, e2e_steps)
        out.update({"e2e_ms": ms_e, "e2e_steps": e2e_steps, "h2d_bytes": xh.numel() * 4 + th.numel() * 4, "d2h_bytes": 4})
    assert torch.isfinite(state["loss"]).all(), "non-finite loss"
    gs = m.backward_graph_stats()
    out.update({"flat_numel": int(m._n_flat), "cuda_graph": {"captured": graphed is not None, "replays": graphed.replays if graphed is not None else 0},
                "loss": float(state["loss"]), "norm": float(state["norm"]),
                "backward_graphs": {"captured": gs[0], "replays": gs[1], "eager_calls": gs[2]}})
    return out


def run_train(args, rank, world, local):
    """BASELINE.json configs[2]: TM-Glow cylinder-array training, data-parallel, one NCCL all-reduce of the flat
    gradient per optimizer step.  One "step" = one BPTT block of `tback` time steps at the global batch
    (trainFlowParallel.py:241-303); strong scaling: the global batch is fixed, each rank takes global_batch / N."""
    assert torch.cuda.is_available()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        with stdout_to_stderr():        # NCCL prints its version banner on stdout while the communicator comes up
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
    clocks = ClockSampler(local)
    clocks.start()
    r = measure_train(args, rank, world, dev, dist, args.steps, args.warmup, e2e_steps=max(2, min(args.steps, 5)))
    clk = clocks.stop()
    ms, loss, norm, launches = r["ms"], r["loss"], r["norm"], r["launches"]
    GB, tb = args.global_batch, args.tback
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            sec, threads, _, kind = cpu_port_train_steps(args.ref_train_batch, tb, 1)
            cpu = {"value": 1.0 / sec, "unit": "steps/s", "cores": threads, "kind": kind,
                   "sample": "1 optimizer step at batch %d x %d time steps through the reference training arithmetic with autograd (%.1f s): "
                             "%.2f HF snapshots/s vs %.0f here" % (args.ref_train_batch, tb, sec, args.ref_train_batch * tb / sec,
                                                                  GB * tb * args.steps / (ms * 1e-3)),
                   "hf_snapshots_per_sec": args.ref_train_batch * tb / sec}
        line = {"metric": "train_steps_per_sec", "value": args.steps / (ms * 1e-3), "unit": "steps/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": DTYPES[args.precision] + " in the forward AND the backward (weight / data gradients on tcgen05)",
                "data": "synthetic",
                "config": {"workload": "TM-Glow cylinder-array training (BASELINE.json configs[2]): global batch %d, BPTT block of %d "
                                       "time steps, x[B,T,3,16,16] -> y[B,T,3,64,64], default model, Adam-amsgrad (weight decay 1e-8), grad clip 1.0, "
                                       "loss = the reference's TMGLowLoss (beta=200: pressure-Poisson + divergence residuals, MSE, RMS, entropy; fused "
                                       "CUDA kernel), LSTM states mixed with the initial states after each step" % (GB, tb),
                           "global_batch": GB, "tback": tb, "parallelism": "dp%d, one all-reduce of the flat gradient per step" % world,
                           "l2_policy": "per-step working set (tape + activations, > 3 GB) exceeds the 126 MB L2; no explicit flush"},
                "clocks": clk, "gpu_launches": int(launches), "loss": loss, "grad_norm": norm,
                "e2e": {"value": r["e2e_steps"] / (r["e2e_ms"] * 1e-3), "unit": "steps/s", "h2d_bytes_per_step": r["h2d_bytes"],
                        "d2h_bytes_per_step": r["d2h_bytes"]},
                "cpu_baseline": cpu, "backward_graphs": r["backward_graphs"], "cuda_graph": r["cuda_graph"],
                "allreduce": {"ms_per_step": r.get("allreduce_ms_per_step", 0.0),
                              "share_of_step": r.get("allreduce_ms_per_step", 0.0) / (ms / args.steps),
                              "bytes": 4 * r.get("flat_numel", 0), "note": "CUDA events around dist.all_reduce of the flat gradient on rank 0 (one per optimizer step)"},
                "hf_snapshots_per_sec": GB * tb * args.steps / (ms * 1e-3)}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


SCALED_GEOM = dict(nic=4, h=64, w=128, noc=3, H=128, W=256)          # BASELINE configs[4]: 2x grid resolution
SCALED_BLOCKS = [24, 24, 24]                                          # deeper flow stack (24 steps per level)
SCALED_ALG_FLOP, SCALED_ALG_BYTES = 9.916e9, 77.6e6                   # BASELINE.md section 3, per HF sample


def _dist_setup(world, local):
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        with stdout_to_stderr():
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
    return dev, dist


def _timed_steps(fn, steps, warmup, dev, dist):
    for _ in range(warmup):
        fn()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def run_uq(args, rank, world, local):
    """BASELINE.json configs[3]: cylinder-array sample-parallel uncertainty quantification through the UQ driver
    (tmglow_b200.uq.sample_sequence = the sample loop of TrainFlow.test / modelPred, trainFlowParallel.py:345-367, folded into
    the batch): ONE low-fidelity sequence of T time steps, S stochastic HF samples PER GPU (samples sharded over the ranks by
    global index), LSTM states mixed with the key states every 10 steps, per-time-step mean / variance fields streamed in fp64
    and combined with ONE all-reduce per sequence -- inside the timed region.  One "step" = one sequence."""
    from tmglow_b200 import TMGlow, _lib, uq
    dev, dist = _dist_setup(world, local)
    lib = _lib.load()
    torch.manual_seed(12345); np.random.seed(12345)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        m = TMGlow(TRAIN_GEOM["nic"], TRAIN_GEOM["noc"], [4, 4, 4], [16, 16, 16], **TRAIN_KW)
    perturb_(m, 12346)
    m = m.to(dev).eval()
    m.precision = args.precision
    S, T = args.samples, args.uq_steps
    g = torch.Generator().manual_seed(11)                      # the SAME sequence on every rank
    x_host = torch.randn(T, TRAIN_GEOM["nic"], TRAIN_GEOM["h"], TRAIN_GEOM["w"], generator=g).pin_memory()
    state = {}
    # the seeded initial LSTM states of this rank's samples (global indices): the reference's HOST generator, one per sample
    # (tmGlow.py:481-509) -- seconds of CPU time for 4096 samples, prepared once outside the timed region like a data pipeline
    lo, hi = uq.shard_range(S * world, rank, world)
    key = m.initLSTMStates(uq.sample_seeds(7, 0, lo, hi), [TRAIN_GEOM["H"], TRAIN_GEOM["W"]])

    def seq():
        x_seq = x_host.to(dev, non_blocking=True)              # H2D of the step's LF sequence
        mean, var, ntot, _ = uq.sample_sequence(m, x_seq, S * world, base_seed=7, sequence=0, state_mix_every=10, rank=rank,
                                                world=world, unnormalise=False, key_states=key)
        state["mean"], state["var"], state["n"] = mean, var, ntot
        state["host"] = mean[-1, :, :4, :4].cpu()              # D2H read of a result (forces completion of the reduction)
    clocks = ClockSampler(local)
    clocks.start()
    lib.tmg_launch_count(1)
    ms = _timed_steps(seq, args.steps, max(args.warmup, 1), dev, dist)
    launches = lib.tmg_launch_count(0)
    clk = clocks.stop()
    assert state["n"] == S * world and torch.isfinite(state["mean"]).all() and torch.isfinite(state["var"]).all()
    value = world * S * T * args.steps / (ms * 1e-3)
    if rank == 0:
        pk = peaks()
        line = {"metric": "hf_samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 1), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": DTYPES[args.precision], "data": "synthetic",
                "config": {"workload": "TM-Glow cylinder-array sample-parallel UQ (BASELINE.json configs[3]): 1 LF sequence x[%d,3,16,16] -> "
                                       "%d stochastic HF samples y[3,64,64] per GPU per time step, %d time steps, moments all-reduce per "
                                       "sequence; uq.sample_sequence" % (T, S, T),
                           "samples_per_gpu": S, "time_steps": T, "parallelism": "sample-parallel x%d, one all-reduce of the moment sums per sequence" % world,
                           "l2_policy": "working set per time step (%.1f GB) exceeds the 126 MB L2; no explicit flush" % (S * 1.8e6 / 1e9),
                           "precision": args.precision},
                "clocks": clk, "gpu_launches": int(launches),
                "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": x_host.numel() * 4, "d2h_bytes_per_step": 3 * 16 * 4,
                        "note": "the timed region IS the end-to-end driver call: pinned-host LF sequence in, moments reduced, result read back"},
                "whole_path": {"alg_tflops": value * 1.072e9 / 1e12 / world, "alg_gbs": value * 6.9e6 / 1e9 / world, "per": "GPU",
                               "frac_hbm": value * 6.9e6 / 1e9 / world / pk["hbm"], "frac_tensor": value * 1.072e9 / 1e12 / world / pk["tf_sust"]}}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def run_scaled(args, rank, world, local):
    """BASELINE.json configs[4]: scaled TM-Glow (24 flow steps per level, HF grid 128x256), sampling + training sweep.
    The spec's "bf16 coupling nets" are served by the single-pass fp16-operand mode `f16` (same 8-bit-exponent-free 11-bit
    significand arithmetic class as bf16's 8 bits, fp32 accumulate; its own tolerance, tests/test_gpu_parity.py); the headline
    value is the fp32-grade f16x3 mode, `fast_mode` reports f16."""
    from tmglow_b200 import TMGlow, FlatAdam, _lib, train as T
    dev, dist = _dist_setup(world, local)
    lib = _lib.load()
    torch.manual_seed(12345); np.random.seed(12345)
    import contextlib, io, types
    with contextlib.redirect_stdout(io.StringIO()):
        m = TMGlow(SCALED_GEOM["nic"], SCALED_GEOM["noc"], [4, 4, 4], SCALED_BLOCKS, **MODEL_KW)
    perturb_(m, 12346)
    m = m.to(dev).eval()
    S = args.samples
    g = torch.Generator().manual_seed(100 + rank)
    x_host = torch.randn(1, SCALED_GEOM["nic"], SCALED_GEOM["h"], SCALED_GEOM["w"], generator=g).pin_memory()
    x_stage = x_host.to(dev)
    h0 = m.initLSTMStates(torch.arange(S) + 1000 * rank, [SCALED_GEOM["H"], SCALED_GEOM["W"]])
    res = {}
    st = {"h": h0}

    def step():
        x_stage.copy_(x_host, non_blocking=True)
        y, ld, st["h"] = m.sample(x_stage.expand(S, -1, -1, -1), st["h"])
        st["ld"] = ld
    clocks = ClockSampler(local)
    clocks.start()
    for prec in (args.precision, "f16"):
        m.precision = prec
        st["h"] = h0
        lib.tmg_launch_count(1)
        ms = _timed_steps(step, args.steps, max(args.warmup, 3), dev, dist)
        res[prec] = (world * S * args.steps / (ms * 1e-3), ms / args.steps, lib.tmg_launch_count(0))
        assert torch.isfinite(st["ld"]).all()
    clk = clocks.stop()
    # training leg: data-parallel BPTT block at a bounded global batch (tape: ~30 MB per sample and time step)
    train = None
    if not args.no_train:
        try:
            del st, h0
            torch.cuda.empty_cache()
            m.train(); m.precision = args.precision
            GB, tb = args.scaled_train_batch, args.tback
            Bl = GB // world
            gt = torch.Generator().manual_seed(7 + rank)
            xb = torch.randn(Bl, tb, SCALED_GEOM["nic"], SCALED_GEOM["h"], SCALED_GEOM["w"], generator=gt).to(dev)
            tg = torch.randn(Bl, tb, SCALED_GEOM["noc"], SCALED_GEOM["H"], SCALED_GEOM["W"], generator=gt).to(dev)
            from tmglow_b200.loss import TMGLowLoss, target_statistics
            crit = TMGLowLoss(types.SimpleNamespace(beta=200.0, dx=2.0 / 64, dy=2.0 / 64), m).to(dev)
            tm, tr = target_statistics(tg)
            hk = m.initLSTMStates(torch.arange(Bl) + 1000 * rank, [SCALED_GEOM["H"], SCALED_GEOM["W"]])
            opt = FlatAdam(m, lr=1e-4, amsgrad=True)
            gr = T.GraphedTrainBlock(m, opt, crit, xb, tg, tm, tr, hk, max_norm=1.0, weight_decay=1e-8)
            ms_t = _timed_steps(lambda: gr.step(), max(args.steps, 3), 1, dev, dist)
            nst = max(args.steps, 3)
            train = {"metric": "train_steps_per_sec", "value": nst / (ms_t * 1e-3), "ms_per_step": ms_t / nst, "global_batch": GB,
                     "tback": tb, "hf_snapshots_per_sec": GB * tb * nst / (ms_t * 1e-3), "loss": float(gr.loss), "cuda_graph": True}
        except Exception as ex:
            train = {"error": repr(ex)[:300]}
    if rank == 0:
        pk = peaks()
        v, msps, launches = res[args.precision]
        line = {"metric": "hf_samples_per_sec", "value": v, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": msps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": DTYPES[args.precision], "data": "synthetic",
                "config": {"workload": "Scaled TM-Glow (BASELINE.json configs[4]): glow_blocks [24,24,24], x[1,4,64,128] -> %d stochastic HF "
                                       "samples y[3,128,256] per GPU per step" % S, "samples_per_gpu_per_step": S,
                           "parallelism": "sample-parallel x%d, no collective" % world, "precision": args.precision,
                           "l2_policy": "working set per step (%.1f GB) exceeds the 126 MB L2; no explicit flush" % (S * 20e6 / 1e9)},
                "clocks": clk, "gpu_launches": int(launches),
                "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": x_host.numel() * 4, "d2h_bytes_per_step": 0,
                        "note": "LF snapshot copied from pinned host memory every step; samples stay on the device (a UQ driver reduces them there)"},
                "fast_mode": {"precision": "f16", "value": res["f16"][0], "unit": "samples/s",
                              "note": "single-pass fp16 operands in place of the spec's bf16 coupling nets; own tolerance"},
                "whole_path": {"alg_tflops": v * SCALED_ALG_FLOP / 1e12 / world, "alg_gbs": v * SCALED_ALG_BYTES / 1e9 / world, "per": "GPU",
                               "frac_hbm": v * SCALED_ALG_BYTES / 1e9 / world / pk["hbm"],
                               "frac_tensor": v * SCALED_ALG_FLOP / 1e12 / world / pk["tf_sust"]},
                "train": train}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()



def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--samples", type=int, default=4096, help="stochastic HF samples per GPU per step (SURVEY 8d C2: 256/1024/4096)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="f16x3", choices=["fp32", "tf32x3", "tf32", "f16x3", "f16"],
                    help="f16x3 (default): tcgen05 with the fp16 hi+lo operand split, fp32-grade (same tolerance as fp32)")
    ap.add_argument("--ref-batch", type=int, default=0, help="batch of the CPU reference arm; 0 = where samples/s saturates (sweep 16/64/256)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="sample", choices=["sample", "train", "uq", "scaled"],
                    help="sample: HF samples/s (configs[1], the default line); train: train steps/s (configs[2]); uq: cylinder "
                         "sample-parallel UQ through uq.sample_sequence (configs[3]); scaled: 24 steps/level at 128x256 (configs[4])")
    ap.add_argument("--uq-steps", type=int, default=40, help="time steps of the LF sequence of --workload uq (trainFlowParallel.py:345)")
    ap.add_argument("--scaled-train-batch", type=int, default=16)
    ap.add_argument("--global-batch", type=int, default=64)
    ap.add_argument("--tback", type=int, default=10)
    ap.add_argument("--ref-train-batch", type=int, default=4, help="batch of the CPU training baseline (bounded sample)")
    ap.add_argument("--no-latency", action="store_true", help="skip the small-batch eager vs CUDA-graph latency leg")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle check of the timed configuration")
    ap.add_argument("--no-train", action="store_true", help="skip the short training measurement of the default line")
    ap.add_argument("--train-steps", type=int, default=10, help="timed optimizer steps of the training legs of the default line")
    ap.add_argument("--no-graph", action="store_true", help="training: launch eagerly instead of replaying the captured CUDA graphs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if args.workload == "train":
            run_reference_train(args, rank, world)
        else:
            run_reference(args, rank, world)
        return
    if args.workload == "train":
        run_train(args, rank, world, local)
        return
    if args.workload == "uq":
        run_uq(args, rank, world, local)
        return
    if args.workload == "scaled":
        if args.samples == 4096:
            args.samples = 512                       # 4x the pixels and 1.5x the steps of configs[1]
        run_scaled(args, rank, world, local)
        return

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        with stdout_to_stderr():        # NCCL prints its version banner on stdout while the communicator comes up
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()

    from tmglow_b200 import _lib
    lib = _lib.load()
    model = build_model()
    model = model.to(dev)
    model.precision = args.precision
    S, K, W = args.samples, args.steps, args.warmup

    g = torch.Generator().manual_seed(100 + rank)
    x_host = torch.randn(1, GEOM["nic"], GEOM["h"], GEOM["w"], generator=g).pin_memory()
    # ONE LF input, S stochastic samples: the batch-expanded view is what a UQ caller passes (the model detects
    # the zero batch stride, runs the encoder once and shares the conditioning maps)
    x_dev = x_host.to(dev).expand(S, -1, -1, -1)
    h0 = model.initLSTMStates(torch.arange(S) + 1000 * rank, [GEOM["H"], GEOM["W"]])
    torch.manual_seed(777 + rank)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if dist is None:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- device-resident run: inputs already in HBM when the timed region starts
    h = h0
    for _ in range(W):
        y, ld, h = model.sample(x_dev, h)
    clocks = ClockSampler(local)
    barrier()
    clocks.start()
    lib.tmg_launch_count(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        y, ld, h = model.sample(x_dev, h)
    e1.record()
    barrier()
    launches = lib.tmg_launch_count(0)
    ms = max_over_ranks(e0.elapsed_time(e1))
    clk = clocks.stop()
    assert torch.isfinite(y).all(), "non-finite samples"
    value = world * S * K / (ms * 1e-3)

    # ---------------- end to end through the public API with HOST buffers: every step copies the LF snapshot from
    # pinned host memory, samples, and copies the S HF fields + log-dets back to pinned host memory.  The D2H copy of
    # step k runs on a copy stream (double-buffered host side) while step k+1 computes, as a production UQ loop would.
    y_host = [torch.empty((S, GEOM["noc"], GEOM["H"], GEOM["W"]), dtype=torch.float32).pin_memory() for _ in range(2)]
    ld_host = [torch.empty(S, dtype=torch.float32).pin_memory() for _ in range(2)]
    x_stage = torch.empty_like(x_host, device=dev)
    copy_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream(dev)

    def e2e_step(k, h):
        x_stage.copy_(x_host, non_blocking=True)                       # H2D of the step's LF input
        y, ld, h = model.sample(x_stage.expand(S, -1, -1, -1), h)
        done = torch.cuda.Event()
        done.record(main_stream)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(done)
            y_host[k & 1].copy_(y, non_blocking=True)                  # D2H of the step's HF samples
            ld_host[k & 1].copy_(ld, non_blocking=True)
            y.record_stream(copy_stream); ld.record_stream(copy_stream)
        return h

    h = h0
    for k in range(2):
        h = e2e_step(k, h)
    main_stream.wait_stream(copy_stream)
    barrier()
    e0.record()
    for k in range(K):
        h = e2e_step(k, h)
    main_stream.wait_stream(copy_stream)                               # the last copy is inside the timed region
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    assert torch.isfinite(y_host[(K - 1) & 1]).all()
    e2e = {"value": world * S * K / (ms_e2e * 1e-3), "unit": "samples/s",
           "h2d_bytes_per_step": x_host.numel() * 4, "d2h_bytes_per_step": (y_host[0].numel() + ld_host[0].numel()) * 4}
    # the device->host link of this box, measured alone (one copy of the step's result into the pinned buffer): the copy of
    # step k overlaps the compute of step k + 1, so a step costs max(compute, copy) -- on a box whose link is slower than
    # d2h_bytes / ms_per_step the end-to-end number is the link's, not the kernels'
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    c0.record()
    for _ in range(2):
        y_host[0].copy_(y, non_blocking=True)
    c1.record()
    torch.cuda.synchronize(dev)
    d2h_ms = c0.elapsed_time(c1) / 2
    e2e["d2h_alone_ms"] = d2h_ms
    e2e["d2h_gb_per_s"] = y_host[0].numel() * 4 / (d2h_ms * 1e-3) / 1e9
    e2e["bound"] = "d2h link" if d2h_ms > ms / K else "compute"

    # ---------------- parity of what was just timed (rank 0, outside the timed region): the same S samples of the same LF input
    # once more with EXPLICIT noise through reconstruct(), a handful of them against the pinned oracle on the CPU
    # (fields 2e-4 abs, log-dets 1e-5 rel: the fp32 tolerance of the test-suite); the sticky fp16-overflow flag must be clear
    parity = None
    if rank == 0 and not args.no_parity:
        from oracle import tmglow_oracle as O
        ocfg = O.OracleConfig.from_dict(model._cfg_dict)
        sd_cpu = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
        gp = torch.Generator(device=dev).manual_seed(4242)
        eps_p = [torch.randn(sh, generator=gp, device=dev) for sh in model.latent_shapes(S, GEOM["H"], GEOM["W"])]
        y_p, ld_p, _ = model.reconstruct(x_dev, h0, eps_p)
        pick = torch.tensor(sorted({0, 1, S // 3, S // 2, S - 2, S - 1}))
        with torch.no_grad():
            y_o, ld_o, _ = O.reconstruct(sd_cpu, ocfg, x_host.expand(len(pick), -1, -1, -1).contiguous(),
                                         [(a[pick].cpu(), c[pick].cpu()) for a, c in h0], [e[pick.to(dev)].cpu() for e in eps_p])
        e_y = (y_p[pick.to(dev)].cpu() - y_o).abs().max().item()
        e_l = ((ld_p[pick.to(dev)].cpu() - ld_o).abs() / ld_o.abs()).max().item()
        ovf = model.activation_overflow()
        parity = {"checked": True, "samples": [int(i) for i in pick], "of": S, "max_abs_y": e_y, "max_rel_log_det": e_l,
                  "tolerance": {"y_abs": 2e-4, "log_det_rel": 1e-5}, "fp16_overflow_flag": bool(ovf)}
        assert e_y < 2e-4 and e_l < 1e-5 and not ovf, "bench parity check failed: %r" % (parity,)
        del eps_p, y_p, ld_p

    # ---------------- the single-pass fp16 mode, reported separately with its own tolerance (tests/test_gpu_parity.py)
    fast = None
    if args.precision == "f16x3":
        model.precision = "f16"
        for _ in range(3):
            y, ld, h = model.sample(x_dev, h)
        barrier()
        e0.record()
        for _ in range(K):
            y, ld, h = model.sample(x_dev, h)
        e1.record()
        barrier()
        ms_f = max_over_ranks(e0.elapsed_time(e1))
        fast = {"precision": "f16", "value": world * S * K / (ms_f * 1e-3), "unit": "samples/s",
                "tolerance": "5e-2 abs on fields, 1e-3 rel on log_det (fp32-grade modes: 2e-4 / 1e-5)"}
        model.precision = args.precision

    # ---------------- small-batch latency: 64 DISTINCT LF inputs per call, eager launches vs one CUDA-graph replay
    # (uq.GraphedSampler: ~100 launches + the Python work of a call become one cudaGraphLaunch)
    latency = None
    if rank == 0 and not args.no_latency:
        from tmglow_b200 import uq
        Bq = 64
        xq = torch.randn(Bq, GEOM["nic"], GEOM["h"], GEOM["w"], generator=torch.Generator().manual_seed(9)).to(dev)
        hq = model.initLSTMStates(torch.arange(Bq), [GEOM["H"], GEOM["W"]])
        stq = {"h": hq}

        def eager_call():
            yq, lq, stq["h"] = model.sample(xq, stq["h"])
        for _ in range(5):
            eager_call()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(30):
            eager_call()
        torch.cuda.synchronize()
        eager_ms = (time.perf_counter() - t0) / 30 * 1e3
        gsamp = uq.GraphedSampler(model, xq, hq)
        for _ in range(5):
            gsamp.sample()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(30):
            gsamp.sample()
        torch.cuda.synchronize()
        graph_ms = (time.perf_counter() - t0) / 30 * 1e3
        latency = {"batch": Bq, "inputs": "distinct", "eager_ms_per_call": eager_ms, "graph_ms_per_call": graph_ms,
                   "graph_replays": gsamp.replays, "samples_per_sec_graph": Bq / (graph_ms * 1e-3),
                   "note": "wall clock per sample() call incl. host work, 30 calls after 5 warm-ups, synchronised at both ends"}
        del gsamp, xq, hq, stq

    # ---------------- per-kernel-class device times (CUDA events on the launching stream)
    pk = peaks()
    roof, roof_tensor, classes = None, None, []
    if rank == 0:
        lib.tmg_profile_enable(1)
        nprof = 2
        for _ in range(nprof):
            y, ld, h = model.sample(x_dev, h)
        torch.cuda.synchronize()
        import ctypes as C
        tot = 0.0
        for t in range(lib.tmg_profile_classes()):
            msv, n, fl, by = C.c_double(), C.c_int64(), C.c_double(), C.c_double()
            _lib.check(lib.tmg_profile_query(t, C.byref(msv), C.byref(n), C.byref(fl), C.byref(by)))
            if n.value:
                classes.append({"class": lib.tmg_profile_class_name(t).decode(), "ms_per_step": msv.value / nprof,
                                "launches_per_step": n.value // nprof,
                                "tflops": fl.value / (msv.value * 1e-3) / 1e12 if msv.value else 0.0,
                                "gbs": by.value / (msv.value * 1e-3) / 1e9 if msv.value else 0.0,
                                "alg_flops_per_launch": fl.value / n.value, "alg_bytes_per_launch": by.value / n.value})
                tot += msv.value / nprof
        lib.tmg_profile_enable(0)
        for c in classes:
            c["share"] = c["ms_per_step"] / tot if tot else 0.0
        if classes:
            top = max(classes, key=lambda c: c["ms_per_step"])
            traffic = None
            tpath = os.path.join(ROOT, "profiles", "traffic.json")      # dram bytes per launch from the ncu --set full capture
            if os.path.exists(tpath):
                traffic = json.load(open(tpath)).get(top["class"])
            if top["class"].startswith("conv"):
                roof = {"kernel": top["class"], "bound": "tensor", "achieved": top["tflops"], "peak": pk["tf_sust"],
                        "unit": "TFLOP/s", "frac": top["tflops"] / pk["tf_sust"], "traffic": traffic,
                        "peak_source": pk["source"] + " (bf16 cuBLAS sustained)", "share_of_step": top["share"],
                        "avg_launch_ms": top["ms_per_step"] / top["launches_per_step"]}
            else:
                roof = {"kernel": top["class"], "bound": "hbm", "achieved": top["gbs"], "peak": pk["hbm"],
                        "unit": "GB/s", "frac": top["gbs"] / pk["hbm"], "traffic": traffic,
                        "peak_source": pk["source"], "share_of_step": top["share"],
                        "avg_launch_ms": top["ms_per_step"] / top["launches_per_step"],
                        "note": "algorithmic bytes = 4*px*(2C+cond) per flow step (SURVEY 8d, un-hoisted definition) x the steps a launch "
                                "runs; flow_level_resident = one launch per level with all 15 plain steps, the state resident on the SM: its "
                                "DRAM traffic (ncu, level-0 launch: 761 MB for 4096 samples = the state read once + written once) is 37x "
                                "below the algorithmic bytes of the 15 steps it replaces (28.2 GB); achieved / avg_launch_ms average over the "
                                "launches of the class in a step"}
            gate = [c for c in classes if c["class"] == "conv_lstm_gates"]
            if gate:
                x3 = args.precision in ("f16x3", "tf32x3")
                roof_tensor = {"kernel": "conv_lstm_gates", "bound": "tensor", "achieved": gate[0]["tflops"],
                               "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": gate[0]["tflops"] / pk["tf_sust"],
                               "executed_tflops": gate[0]["tflops"] * (3 if x3 else 1),
                               "note": "achieved = algorithmic 2*M*N*K flops; the hi/lo split executes 3 MMAs per algorithmic one"}

    # ---------------- CPU baseline (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, smp, cb, ncalls, el, sweep = cpu_samples_per_s(args.cpu_seconds)
        cpu = {"value": v, "unit": "samples/s", "cores": smp.threads, "kind": smp.kind,
               "sample": "%d x %s of %d HF samples (%.1f s) on %d host threads; samples/s by batch: %s" % (
                   ncalls, "the unmodified reference TMGlow.sample (oracle/_ref)" if smp.kind == "reference" else "oracle.sample()",
                   cb, el, smp.threads, {k: round(x_, 1) for k, x_ in sweep.items()})}

    # ---------------- the other half of BASELINE.json's metric: data-parallel train steps/s (configs[2]) -- the path with
    # the collective: one all-reduce of the flat gradient per optimizer step.  Strong scaling (global batch 64 split over the
    # ranks, BASELINE configs[2]) and, for N > 1, weak scaling (64 per GPU); >= 10 timed optimizer steps each.
    train, train_weak = None, None
    if not args.no_train:
        del y, ld, h
        model = None
        torch.cuda.empty_cache()

        def train_leg(global_batch, scaling):
            a2 = argparse.Namespace(**vars(args))
            a2.global_batch = global_batch
            try:
                r_t = measure_train(a2, rank, world, dev, dist, steps=args.train_steps, warmup=2)
                ms_t = r_t["ms"]
                return {"metric": "train_steps_per_sec", "value": args.train_steps / (ms_t * 1e-3), "unit": "steps/s",
                        "ms_per_step": ms_t / args.train_steps, "steps": args.train_steps, "global_batch": global_batch, "tback": args.tback,
                        "scaling": scaling, "hf_snapshots_per_sec": global_batch * args.tback * args.train_steps / (ms_t * 1e-3),
                        "parallelism": "dp%d, one all-reduce of the flat gradient per step" % world,
                        "allreduce_ms_per_step": r_t.get("allreduce_ms_per_step", 0.0),
                        "allreduce_share_of_step": r_t.get("allreduce_ms_per_step", 0.0) / (ms_t / args.train_steps),
                        "allreduce_bytes": 4 * r_t.get("flat_numel", 0), "cuda_graph": r_t["cuda_graph"],
                        "gpu_launches": r_t["launches"], "loss": r_t["loss"],
                        "workload": "TM-Glow cylinder-array training (configs[2]); see bench.py --workload train"}
            except Exception as ex:        # the sampling line must not be lost to a training failure
                return {"error": repr(ex)[:300]}
        train = train_leg(args.global_batch, "strong")
        if world > 1:
            torch.cuda.empty_cache()
            train_weak = train_leg(args.global_batch * world, "weak")

    if rank == 0:
        line = {
            "metric": "hf_samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": DTYPES[args.precision],
            "data": "synthetic", "config": dict(workload_config(S, world), precision=args.precision),
            "clocks": clk, "e2e": e2e, "gpu_launches": int(launches),
            "parity_checked": bool(parity and parity["checked"]), "parity": parity, "latency_small_batch": latency,
            "roofline": roof, "roofline_tensor": roof_tensor, "cpu_baseline": cpu, "fast_mode": fast, "train": train, "train_weak": train_weak,
            "whole_path": {"alg_tflops": value * ALG_FLOP_PER_SAMPLE / 1e12 / world,
                           "alg_gbs": value * ALG_BYTES_PER_SAMPLE / 1e9 / world, "per": "GPU"},
            "kernel_classes": classes,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
