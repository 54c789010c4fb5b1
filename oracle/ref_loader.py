"""Imports the UNMODIFIED reference package (``oracle/_ref/tmglow``, made by ``oracle/make_ref.py``; in the build container
also ``/root/reference/tmglow``) for the reference arm of ``bench.py`` and for tests.  TEST / BENCH INFRASTRUCTURE ONLY:
nothing under ``deep-turbulence_b200/`` imports this module.

The reference is run from its source tree (``main.py`` puts ``tmglow/`` on ``sys.path`` and imports ``nn.tmGlow``), so its
top-level package names are ``nn``, ``pc`` and ``utils``.  Two harness-side accommodations, neither edits a reference file:
  * ``utils/viz.py`` imports matplotlib (absent in this image, plotting only): an inert stand-in is registered;
  * ``GaussianDiag.__init__`` clamps a ``chunk`` view in place (``nn/modules/flowUtils.py:163``), which current PyTorch
    refuses under autograd: ``grad_shim()`` replaces the constructor by the out-of-place form for the TRAINING arm only
    (SURVEY.md 8c caveat 1); inference runs the class as it is.
"""
import os
import sys
import types
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
CANDIDATES = (os.path.join(HERE, "_ref", "tmglow"), "/root/reference/tmglow")


def ref_root():
    for c in CANDIDATES:
        if os.path.exists(os.path.join(c, "nn", "tmGlow.py")):
            return c
    return None


def _matplotlib_stand_in():
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.gridspec", "matplotlib.colors", "matplotlib.ticker"):
        if name not in sys.modules:
            try:
                __import__(name)
            except ImportError:
                sys.modules[name] = types.ModuleType(name)
    mpl = sys.modules["matplotlib"]
    if not hasattr(mpl, "use"):
        class _Anything(dict):
            def __call__(self, *a, **k):
                return self

            def __getattr__(self, k):
                return self
        mpl.__path__ = []
        mpl.use = lambda *a, **k: None
        mpl.rcParams = _Anything()
        mpl.rc = lambda *a, **k: None
        mpl.pyplot = sys.modules["matplotlib.pyplot"]
        def _attr(k):
            if k.startswith("__"):              # inspect / importlib probe __file__, __path__, ...: those stay absent
                raise AttributeError(k)
            return _Anything()
        for sub in ("pyplot", "gridspec", "colors", "ticker"):
            sys.modules["matplotlib." + sub].__getattr__ = _attr


def load(trainer=False):
    """Returns a namespace with the reference's ``TMGlow`` (and, with ``trainer=True``, ``TMGLowLoss``), or raises
    ``ImportError`` when no copy of the reference is available."""
    root = ref_root()
    if root is None:
        raise ImportError("reference package not found (run `python oracle/make_ref.py` in the build container)")
    if root not in sys.path:
        sys.path.insert(0, root)
    for name in ("nn", "pc", "utils"):          # a foreign top-level module of the same name would shadow the reference's
        mod = sys.modules.get(name)
        if mod is not None and not str(getattr(mod, "__file__", "") or "").startswith(root):
            raise ImportError("module %r is already imported from %s" % (name, getattr(mod, "__file__", None)))
    ns = types.SimpleNamespace(root=root)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from nn.tmGlow import TMGlow
        ns.TMGlow = TMGlow
        if trainer:
            _matplotlib_stand_in()
            from nn.trainFlowParallel import TMGLowLoss
            ns.TMGLowLoss = TMGLowLoss
    return ns


def grad_shim():
    """Out-of-place clamp in ``GaussianDiag.__init__`` (flowUtils.py:157-163) so that ``sample()`` runs under autograd."""
    import math
    from nn.modules import flowUtils as FU

    def _init(self, mean, log_stddev):
        self.mean = mean
        self.log_stddev = log_stddev.clamp(min=-10., max=math.log(5.))
    if getattr(FU.GaussianDiag.__init__, "_tmg_shim", False):
        return
    _init._tmg_shim = True
    FU.GaussianDiag.__init__ = _init
