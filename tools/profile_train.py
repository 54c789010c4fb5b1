"""Per-kernel device time of one data-parallel training step (bench.py --workload train) with torch.profiler (CUPTI):
cheap (one pass, no replay) -- the ncu launch list of a 16 000-launch step takes minutes.
    python tools/profile_train.py [--global-batch 64] [--tback 10] [--precision f16x3] > gpurun_out/train_kernels.txt
"""
import argparse
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "deep-turbulence_b200"))
import torch
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--global-batch", type=int, default=64)
ap.add_argument("--tback", type=int, default=10)
ap.add_argument("--precision", default="f16x3")
args = ap.parse_args()
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
from torch.profiler import profile, ProfilerActivity
bench.measure_train(args, 0, 1, dev, None, steps=1, warmup=2)          # warm everything (allocator, derived weights)
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    r = bench.measure_train(args, 0, 1, dev, None, steps=1, warmup=1)
    ms = r["ms"]
    torch.cuda.synchronize()
agg = collections.OrderedDict()
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        n, t = agg.get(ev.name, (0, 0.0))
        agg[ev.name] = (n + 1, t + ev.device_time)
tot = sum(t for _, t in agg.values())
print("# torch.profiler (CUPTI) kernel times of measure_train(steps=1, warmup=1) -> 2 optimizer steps; step time %.1f ms" % ms)
print("# %d device activities, %.1f ms summed" % (sum(n for n, _ in agg.values()), tot / 1e3))
print("%-90s %8s %12s %7s" % ("kernel", "launches", "total_us", "share"))
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
    print("%-90s %8d %12.1f %6.1f%%" % (name[:90], n, t, 100.0 * t / tot))
