"""Import the reference's own ``pipeline.py`` under Python 3 (test infrastructure).

Only usable where ``/root/reference`` exists (the build container, NOT the GPU
box).  Used by ``scripts/make_golden.py`` to freeze golden vectors and by the
CPU tests (skipped when the reference tree is absent) to pin
``oracle.weightmap_oracle`` against the real reference code.

Recipe: ``pipeline.py`` is py3-syntax-clean; its two
missing imports (``skimage.transform``, ``matplotlib.pyplot``,
``pipeline.py:29,34``) are stubbed and the two py2 builtins it relies on
(``zip`` returning a list at ``pipeline.py:538``; ``xrange``) are injected as
module globals.  Nothing in the read-only tree is modified.
"""
import os
import sys
import types
import warnings

REFERENCE_ROOT = os.environ.get("SEQUITR_REFERENCE", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "sequitr", "pipeline.py"))


_cached = None


def load_reference_pipeline():
    """Return the reference ``pipeline`` module (imported from /root/reference)."""
    global _cached
    if _cached is not None:
        return _cached
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    for name in ("skimage", "skimage.transform", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    st = sys.modules["skimage.transform"]
    if not hasattr(st, "rotate"):
        st.rotate = None
        st.resize = None
    import importlib.util
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        spec = importlib.util.spec_from_file_location(
            "_sequitr_reference_pipeline",
            os.path.join(REFERENCE_ROOT, "sequitr", "pipeline.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    mod.zip = lambda *a: list(zip(*a))
    mod.xrange = range
    _cached = mod
    return mod
