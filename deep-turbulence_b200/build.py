"""Builds libtmglow_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python deep-turbulence_b200/build.py [--force]

Output: deep-turbulence_b200/tmglow_b200/lib/libtmglow_b200.so  (git-ignored, travels with gpurun).
"""
import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "tmglow_b200", "lib")
LIB = os.path.join(OUT_DIR, "libtmglow_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "--use_fast_math=false"]
FLAGS = [f for f in FLAGS if not f.startswith("--use_fast_math")]   # IEEE division / accurate expf on purpose
if os.environ.get("TMG_LV_PROFILE"):                                 # developer build: role cycle counters in flow_level_f16.cu
    FLAGS += ["-DTMG_LV_PROFILE"]
if os.environ.get("TMG_GT_PROFILE"):                                 # developer build: wait-time counters in lstm_gate_f16.cu
    FLAGS += ["-DTMG_GT_PROFILE"]
if os.environ.get("TMG_NVCC_DEFS"):                                  # experiments: extra -D flags (e.g. tile configurations)
    FLAGS += os.environ["TMG_NVCC_DEFS"].split()
if os.environ.get("TMG_MBAR_SLEEP"):                                 # experiment: nanosleep back-off in mbarrier waits
    FLAGS += ["-DTMG_MBAR_SLEEP=" + os.environ["TMG_MBAR_SLEEP"]]


def _newer(src_files, target):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in src_files)


def build(force=False, verbose=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    srcs = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(HERE, "..", "include", "tmglow_b200.h"))
    objs = []

    def compile_one(src):
        obj = os.path.join(OUT_DIR, os.path.basename(src)[:-3] + ".o")
        if force or _newer([src] + hdrs, obj):
            cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
            if verbose:
                print(r.stderr)
        return obj

    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    if force or _newer(objs, LIB):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
