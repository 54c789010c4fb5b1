"""CPU-only checks of the host side: the C-ABI library loads and exports every declared symbol,
the drop-in module keeps the reference's state_dict layout, argument errors, no CPU fallback."""
import ctypes
import json
import os
import re

import pytest
import torch

from conftest import GOLDEN, ROOT, load_golden


def _lib():
    from tmglow_b200 import _lib
    return _lib


def test_library_exports_every_declared_symbol():
    L = _lib()
    hdr = open(os.path.join(ROOT, "include", "tmglow_b200.h")).read()
    declared = set(re.findall(r"\b(tmg_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(L.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), "libtmglow_b200.so does not export " + name
    assert declared == set(L.SIGNATURES), declared ^ set(L.SIGNATURES)
    assert L.load().tmg_version() >= 100


def test_default_state_dict_layout_matches_reference():
    from tmglow_b200 import TMGlow
    lay = json.load(open(os.path.join(GOLDEN, "default_state_dict_layout.json")))
    m = TMGlow(4, 3, [4, 4, 4], [16, 16, 16], cond_features=32, cglow_upscale=2, growth_rate=4,
               init_features=16, rec_features=64)
    mine = [[k, list(v.shape), str(v.dtype)] for k, v in m.state_dict().items()]
    assert mine == lay["entries"]                 # 873 keys, same order, shapes and dtypes
    assert m._num_parameters() == lay["n_parameters"] == 1746573
    assert m.glow_blocks == [16, 16, 16] and m.rec_features == 64
    assert hasattr(m, "encoder") and hasattr(m, "glow") and m.in_mu.shape == (3,)


@pytest.mark.parametrize("name", ["caseA_states", "caseB_up4", "caseC_up1"])
def test_reference_checkpoints_load(name):
    from tmglow_b200 import TMGlow
    g = load_golden(name)
    cfg = json.loads(g["config"])
    m = TMGlow(cfg["in_features"], cfg["out_features"], cfg["enc_blocks"], cfg["glow_blocks"],
               cond_features=cfg["cond_features"], cglow_upscale=cfg["cglow_upscale"], growth_rate=cfg["growth_rate"],
               init_features=cfg["init_features"], rec_features=cfg["rec_features"])
    res = m.load_state_dict(g["state_dict"], strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    for k, v in m.state_dict().items():
        assert torch.equal(v, g["state_dict"][k]), k


def test_checkpoint_with_stale_log_s_old_loads():
    """a checkpoint saved after test() has log_s_old == log_s, which crashes a freshly
    built reference model; the drop-in ignores the cache key."""
    from tmglow_b200 import TMGlow
    g = load_golden("caseA_states")
    cfg = json.loads(g["config"])
    sd = dict(g["state_dict"])
    for k in list(sd):
        if k.endswith("log_s_old"):
            sd[k] = sd[k[:-4]].clone()
    m = TMGlow(cfg["in_features"], cfg["out_features"], cfg["enc_blocks"], cfg["glow_blocks"],
               cond_features=cfg["cond_features"], cglow_upscale=cfg["cglow_upscale"], growth_rate=cfg["growth_rate"],
               init_features=cfg["init_features"], rec_features=cfg["rec_features"])
    m.load_state_dict(sd)


def test_init_lstm_states_matches_reference():
    from tmglow_b200 import TMGlow
    g = load_golden("caseA_states")
    cfg = json.loads(g["config"])
    m = TMGlow(cfg["in_features"], cfg["out_features"], cfg["enc_blocks"], cfg["glow_blocks"],
               cond_features=cfg["cond_features"], cglow_upscale=cfg["cglow_upscale"], growth_rate=cfg["growth_rate"],
               init_features=cfg["init_features"], rec_features=cfg["rec_features"])
    B = g["x"].shape[0]
    st = m.initLSTMStates(torch.arange(B) + 7, list(g["y"].shape[2:]))
    for (h, c), (hr, cr) in zip(st, g["h_in"]):
        assert torch.equal(h, hr) and torch.equal(c, cr)


def test_no_cpu_fallback_and_bad_config():
    from tmglow_b200 import TMGlow
    m = TMGlow(4, 3, [2, 2], [3, 3], cond_features=8, cglow_upscale=2, growth_rate=4, init_features=8, rec_features=8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.sample(torch.zeros(1, 4, 8, 16))
    with pytest.raises(AssertionError):
        TMGlow(4, 3, [2, 2], [3], cond_features=8)
    L = _lib()
    assert L.load().tmg_workspace_bytes(None, 1, 8, 8) == 0
    # product code never imports the oracle
    for root, _, files in os.walk(os.path.join(ROOT, "deep-turbulence_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert "tmglow_oracle" not in src and "from oracle" not in src and "import oracle" not in src, f


def _run_bench(args):
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py")] + args, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "stdout must carry exactly one JSON line: %r" % lines
    return json.loads(lines[0])


def test_bench_reference_arm_contract_sampling():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours): one JSON line with the contract's keys."""
    d = _run_bench(["--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-batch", "2"])
    assert d["impl"] == "reference" and d["metric"] == "hf_samples_per_sec" and d["unit"] == "samples/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1
    # "reference" = the unmodified reference package (oracle/_ref, made by oracle/make_ref.py); "port" only without that copy
    have_ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "tmglow", "nn", "tmGlow.py")) or os.path.exists("/root/reference/tmglow/nn/tmGlow.py")
    assert d["cpu_baseline"]["kind"] == ("reference" if have_ref else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["vs_baseline"] is None


def test_bench_reference_arm_contract_training():
    d = _run_bench(["--impl", "reference", "--workload", "train", "--steps", "1", "--warmup", "0", "--ref-train-batch", "1",
                    "--tback", "2"])
    assert d["impl"] == "reference" and d["metric"] == "train_steps_per_sec" and d["unit"] == "steps/s"
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] in ("reference", "port") and d["e2e"]["value"] == d["value"]
    assert d["config"]["tback"] == 2 and d["config"]["global_batch"] == 1 and d["hf_snapshots_per_sec"] > 0


def test_trainable_mask_and_scratch_pool_cpu():
    """Host-side helpers of the training path: the trainable mask marks exactly the nn.Parameters inside the flat buffer
    (what the reference's model.parameters() yields); the scratch pool hands back the same storage for the same request."""
    from tmglow_b200 import TMGlow
    m = TMGlow(4, 3, [2, 2], [3, 3], cond_features=8, cglow_upscale=2, growth_rate=4, init_features=8, rec_features=8)
    mask = m.trainable_mask()
    flat = m.flat_parameters()
    assert mask.shape == flat.shape and set(mask.unique().tolist()) <= {0.0, 1.0}
    assert int(mask.sum()) == sum(p.numel() for p in m.parameters()) == m._num_parameters()
    table = {name: (off, numel) for name, off, numel, shape in m._table}
    pnames = {n for n, _ in m.named_parameters()}
    for name, (off, numel) in table.items():
        assert float(mask[off:off + numel].min()) == float(mask[off:off + numel].max()) == (1.0 if name in pnames else 0.0), name
    a = m._scratch("g_y", (2, 3, 4, 4), torch.float32, torch.device("cpu"))
    b = m._scratch("g_y", (2, 3, 4, 4), torch.float32, torch.device("cpu"))
    c = m._scratch("g_y", (3, 3, 4, 4), torch.float32, torch.device("cpu"))
    assert a.data_ptr() == b.data_ptr() and c.shape[0] == 3
    assert len([k for k in m._scratch_pool if k[0] == "g_y"]) == 1        # one buffer per name is kept alive


def test_replica_owns_its_handles_and_resolves_its_own_leaves():
    """nn.DataParallel.replicate support (reference caller: utils/parallel.py:150-169, one thread per replica): a replica
    gets its own library handles / flat buffer / workspace, and the flat buffer is assembled from the tensors of the
    REPLICA's module tree (the per-device copies), never from the original's."""
    import torch
    from tmglow_b200 import TMGlow
    m = TMGlow(2, 3, [1], [2], cond_features=4, cglow_upscale=1, growth_rate=2, init_features=4, rec_features=4)
    # what torch.nn.parallel.replicate does, for one replica on the CPU: shallow replicas of every module, re-wired
    mods = list(m.modules())
    idx = {mod: i for i, mod in enumerate(mods)}
    reps = [mod._replicate_for_data_parallel() for mod in mods]
    for i, mod in enumerate(mods):
        for key, child in mod._modules.items():
            setattr(reps[i], key, reps[idx[child]])
        for key, p in mod._parameters.items():
            setattr(reps[i], key, p.detach().clone() + 1.0)          # a distinguishable "broadcast copy"
        for key, b in mod._buffers.items():
            setattr(reps[i], key, b.detach().clone())
    r = reps[0]
    assert r._handles is not m._handles and r._handles == {} and r._flat is None and r._ws == {}
    name0 = next(n for n, _ in m.named_parameters())
    i0 = [n for n, _, _, _ in m._table].index(name0)
    assert torch.equal(r._leaf_tensor(i0), m._leaf_tensor(i0) + 1.0)
    assert r._sync_flat(torch.device("cpu")) is True
    off, numel = m._table[i0][1], m._table[i0][2]
    assert torch.equal(r._flat[off:off + numel], (m._leaf_tensor(i0) + 1.0).reshape(-1))
    # the original is untouched: its tensors were not re-bound to the replica's flat buffer
    m._sync_flat(torch.device("cpu"))
    assert m._leaf_tensor(i0).data_ptr() == m._flat.data_ptr() + 4 * off
    assert m._flat.data_ptr() != r._flat.data_ptr()
