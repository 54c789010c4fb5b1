import argparse, collections, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "deep-turbulence_b200"))
import torch, bench
from torch.profiler import profile, ProfilerActivity
args = argparse.Namespace(global_batch=64, tback=10, precision="f16x3")
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
bench.measure_train(args, 0, 1, dev, None, steps=1, warmup=2)
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], with_stack=True) as prof:
    bench.measure_train(args, 0, 1, dev, None, steps=1, warmup=1)
    torch.cuda.synchronize()
cnt = collections.Counter()
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CPU and ("cudaMemcpyAsync" in ev.name or "Memcpy" in ev.name):
        st = [s for s in (ev.stack or []) if "tmglow" in s or "bench" in s or "train" in s]
        cnt[(ev.name, tuple(st[:3]))] += 1
for k, v in cnt.most_common(12):
    print(v, k)
# parents of memcpy runtime calls
cnt2 = collections.Counter()
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CPU and ev.name.startswith("aten::") and any(c.name in ("cudaMemcpyAsync",) for c in ev.cpu_children):
        cnt2[ev.name] += 1
print(cnt2.most_common(10))
