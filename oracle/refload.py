"""TEST INFRASTRUCTURE ONLY - loader that imports the unmodified reference from /root/reference.

Used in the build container by oracle/gen_golden.py (to write tests/golden/*) and by the CPU tests that
cross-check the oracle restatements against the real reference when /root/reference is mounted.
/root/reference does not exist on the GPU box: nothing under `-m gpu`, smoke() or bench.py imports this.

Recipe: the reference's utils.py imports nine packages that are not installed and
cannot be installed offline (Levenshtein, optuna, matplotlib(.pyplot), unidecode, tesserocr, easyocr,
google.cloud(.vision)); empty stub modules are registered for them. `Levenshtein.distance` is bound to the
oracle's own restatement (python-Levenshtein 0.12.0 is not vendored: see oracle/levenshtein.c).
"""
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIPPED_ROOT = os.path.join(_REPO, "baseline", "_ref")   # git-ignored copy made by install() - it travels to the GPU box


def _pick_root():
    for cand in (os.environ.get("QEB_REFERENCE_ROOT"), "/root/reference", SHIPPED_ROOT):
        if cand and os.path.isdir(os.path.join(cand, "models")):
            return cand
    return os.environ.get("QEB_REFERENCE_ROOT", "/root/reference")


REFERENCE_ROOT = _pick_root()


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "models"))


def install(src="/root/reference", dst=SHIPPED_ROOT):
    """Copy the UNMODIFIED reference tree (python sources + its small JSON artifacts; 7 MB) to baseline/_ref so that the
    drop-in trainer tests and the reference arm of bench.py can run it on the GPU box, where /root/reference does not exist.
    The reference has no setup.py / pyproject.toml, so `pip install --target baseline/_ref /root/reference` has nothing to
    build: a byte-for-byte copy is the install. baseline/_ref is git-ignored (never committed), not gpurun-ignored."""
    import shutil

    if not os.path.isdir(os.path.join(src, "models")):
        return None
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    shutil.copytree(src, dst, ignore=shutil.ignore_patterns(".git", "__pycache__", "*.pyc"))
    return dst


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_loaded = False


def load():
    """Put the reference on sys.path behind the stub modules; returns a namespace of its hot-path symbols."""
    global _loaded
    if not available():
        raise RuntimeError(f"reference not mounted at {REFERENCE_ROOT}")
    if not _loaded:
        import wandb  # noqa: F401  (real package first: it needs google.protobuf)
        from . import pyoracle

        _stub("Levenshtein", distance=pyoracle.levenshtein)
        _stub("optuna", TrialPruned=Exception)
        mpl = _stub("matplotlib")
        mpl.pyplot = _stub("matplotlib.pyplot")
        _stub("unidecode", unidecode=lambda s: s)
        _stub("tesserocr")
        _stub("easyocr")
        try:
            import google  # noqa: F401
        except ImportError:
            _stub("google")
        gc = _stub("google.cloud")
        gc.vision = _stub("google.cloud.vision")
        sys.modules["google"].cloud = gc
        if REFERENCE_ROOT not in sys.path:
            sys.path.insert(0, REFERENCE_ROOT)
        _loaded = True
    import importlib

    ns = types.SimpleNamespace()
    ns.model_crnn = importlib.import_module("models.model_crnn")
    ns.model_unet = importlib.import_module("models.model_unet")
    ns.utils = importlib.import_module("utils")
    ns.selection_utils = importlib.import_module("selection_utils")
    ns.transform_helper = importlib.import_module("transform_helper")
    ns.tracking_utils = importlib.import_module("tracking_utils")
    ns.properties = importlib.import_module("properties")
    return ns
