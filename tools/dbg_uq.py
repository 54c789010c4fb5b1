import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "deep-turbulence_b200")]
import torch, numpy as np, contextlib, io
import bench
from tmglow_b200 import TMGlow, _lib, uq
dev = torch.device("cuda:0")
with contextlib.redirect_stdout(io.StringIO()):
    m = TMGlow(3, 3, [4, 4, 4], [16, 16, 16], **bench.TRAIN_KW)
bench.perturb_(m, 1)
m = m.to(dev).eval(); m.precision = "f16x3"
S = 4096
x = torch.randn(1, 3, 16, 16, device=dev)
h = m.initLSTMStates(torch.arange(S), [64, 64])
def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
st = {"h": h}
def samp():
    y, ld, st["h"] = m.sample(x.expand(S, -1, -1, -1), st["h"]); st["y"] = y
print("sample ms", t(samp))
y = st["y"]
print("moments fp64 ms", t(lambda: (y.double().sum(0), (y.double() * y.double()).sum(0))))
print("moments fp32 ms", t(lambda: (y.sum(0), (y * y).sum(0))))
key = h
print("mix ms", t(lambda: uq.mix_states(st["h"], key)))
lib = _lib.load()
lib.tmg_profile_enable(1)
samp(); torch.cuda.synchronize()
import ctypes as C
for tg in range(lib.tmg_profile_classes()):
    msv, n, fl, by = C.c_double(), C.c_int64(), C.c_double(), C.c_double()
    lib.tmg_profile_query(tg, C.byref(msv), C.byref(n), C.byref(fl), C.byref(by))
    if n.value: print(lib.tmg_profile_class_name(tg).decode(), round(msv.value, 2), n.value)
