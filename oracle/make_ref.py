"""Recipe for ``oracle/_ref``: a byte-for-byte copy of the reference's Python package (``/root/reference/tmglow``: ``nn``,
``pc``, ``utils``, MIT licence file beside it), made from the sources where they lie.

TEST / BENCH INFRASTRUCTURE ONLY.  ``oracle/_ref/`` is git-ignored (the history stays free of reference sources) but it is
NOT gpurun-ignored, so the unmodified reference travels to the GPU box, where ``/root/reference`` does not exist:
``bench.py --impl reference`` and the ``cpu_baseline`` leg time the REAL ``TMGlow.sample`` there (``kind: "reference"``).
The product path never imports it.

    python oracle/make_ref.py          # run in the build container; __graft_entry__.build() does it when needed
"""
import os
import shutil
import sys

SRC = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")


def make(force=False):
    """Returns the path of oracle/_ref (existing copy or fresh), or None when the reference is not available."""
    stamp = os.path.join(DST, "tmglow", "nn", "tmGlow.py")
    if os.path.exists(stamp) and not force:
        return DST
    if not os.path.isdir(os.path.join(SRC, "tmglow")):
        return None
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    shutil.copytree(os.path.join(SRC, "tmglow"), os.path.join(DST, "tmglow"),
                    ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    for f in ("LICENSE",):
        if os.path.exists(os.path.join(SRC, f)):
            shutil.copy(os.path.join(SRC, f), os.path.join(DST, f))
    with open(os.path.join(DST, "README"), "w") as fh:
        fh.write("Unmodified copy of zabaras/deep-turbulence tmglow/ made by oracle/make_ref.py (git-ignored).\n")
    return DST


if __name__ == "__main__":
    p = make(force="--force" in sys.argv)
    print(p if p else "reference sources not found at %s" % SRC)
