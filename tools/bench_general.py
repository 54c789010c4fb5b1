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
