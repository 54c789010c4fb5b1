import sys, time, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/deep-turbulence_b200")
import bench
m = bench.build_model().cuda(); m.precision = sys.argv[1] if len(sys.argv) > 1 else "f16x3"
B = 1024
g = torch.Generator().manual_seed(0)
x = torch.randn(B, 4, 32, 64, generator=g).cuda()      # DISTINCT LF inputs: no shared-input fast path
y = torch.randn(B, 3, 64, 128, generator=g).cuda()
h = m.initLSTMStates(torch.arange(B), [64, 128])
for name, fn in (("sample(distinct x)", lambda h: m.sample(x, h)[2]), ("forward(distinct x)", lambda h: m.forward(x, y, h)[2])):
    hh = h
    for _ in range(2): hh = fn(hh)
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): hh = fn(hh)
    e1.record(); torch.cuda.synchronize()
    print(name, m.precision, "%.1f ms/call  %.0f samples/s" % (e0.elapsed_time(e1) / 5, B * 5 / (e0.elapsed_time(e1) * 1e-3)))

# per-kernel-class device times of one un-shared sample() call
import ctypes as C
from tmglow_b200 import _lib
lib = _lib.load()
lib.tmg_profile_enable(1)
hh = m.sample(x, h)[2]
torch.cuda.synchronize()
for t in range(lib.tmg_profile_classes()):
    msv, n, fl, by = C.c_double(), C.c_int64(), C.c_double(), C.c_double()
    lib.tmg_profile_query(t, C.byref(msv), C.byref(n), C.byref(fl), C.byref(by))
    if n.value:
        print("   %-18s %7.3f ms  %3d launches  %7.1f TFLOP/s  %7.1f GB/s" % (lib.tmg_profile_class_name(t).decode(), msv.value, n.value,
              fl.value / (msv.value * 1e-3) / 1e12, by.value / (msv.value * 1e-3) / 1e9))
lib.tmg_profile_enable(0)
