"""
TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Import helper for the shimmed reference copy made by ``oracle/make_ref.py``.
``load()`` returns a namespace with the reference's model classes and
distribution functions, or ``None`` when ``oracle/_ref`` is absent (it is
git-ignored; it exists in the build container and travels to the GPU box).
"""
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_PARENT = os.path.join(HERE, "_ref")
REF_ROOT = os.path.join(REF_PARENT, "BNMTF_ARD")

_cached = None


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "code", "models"))


def _add_paths():
    # Python-2 implicit relative imports of the reference.
    paths = [
        os.path.join(REF_PARENT, "_stubs"),          # matplotlib stub
        REF_PARENT,                                    # `import BNMTF_ARD...`
        os.path.join(REF_ROOT, "code", "models"),     # `from distributions.x import`
        os.path.join(REF_ROOT, "code", "models", "distributions"),  # `import rtnorm`
        os.path.join(REF_ROOT, "code", "cross_validation"),         # `from mask import`
    ]
    try:
        import matplotlib  # noqa: F401  (real one present: no stub needed)
        paths = paths[1:]
    except Exception:
        pass
    for p in paths:
        if p not in sys.path:
            sys.path.append(p)


def load():
    """Return a namespace of reference classes/functions, or None."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        return None
    _add_paths()
    ns = types.SimpleNamespace()
    base = "BNMTF_ARD.code.models."
    for name in ["bnmf_gibbs", "bnmf_vb", "nmf_icm", "nmf_np",
                 "bnmtf_gibbs", "bnmtf_vb", "nmtf_icm", "nmtf_np"]:
        mod = importlib.import_module(base + name)
        setattr(ns, name, getattr(mod, name))
    d = "BNMTF_ARD.code.models.distributions."
    tn = importlib.import_module(d + "truncated_normal")
    tnv = importlib.import_module(d + "truncated_normal_vector")
    gm = importlib.import_module(d + "gamma")
    ex = importlib.import_module(d + "exponential")
    ns.TN_draw, ns.TN_expectation, ns.TN_variance, ns.TN_mode = (
        tn.TN_draw, tn.TN_expectation, tn.TN_variance, tn.TN_mode)
    ns.TN_vector_draw, ns.TN_vector_expectation = tnv.TN_vector_draw, tnv.TN_vector_expectation
    ns.TN_vector_variance, ns.TN_vector_mode = tnv.TN_vector_variance, tnv.TN_vector_mode
    ns.gamma_draw, ns.gamma_expectation = gm.gamma_draw, gm.gamma_expectation
    ns.gamma_expectation_log, ns.gamma_mode = gm.gamma_expectation_log, gm.gamma_mode
    ns.exponential_draw = ex.exponential_draw
    ns.root = REF_ROOT
    _cached = ns
    return ns


def load_cv():
    """Reference cross-validation modules (mask + the four drivers), or None.  The reference
    imports `sklearn.cross_validation.StratifiedKFold(labels, n_folds, shuffle)` (removed from
    sklearn): SURVEY.md 8c item 5 -- alias it onto sklearn.model_selection."""
    if not available():
        return None
    _add_paths()
    if "sklearn.cross_validation" not in sys.modules:
        from sklearn.model_selection import StratifiedKFold as _SKF
        import numpy as _np

        class StratifiedKFold(object):
            def __init__(self, labels, n_folds, shuffle=False):
                self._labels, self._n, self._shuffle = list(labels), n_folds, shuffle

            def __iter__(self):
                skf = _SKF(n_splits=self._n, shuffle=self._shuffle)
                return iter(skf.split(_np.zeros(len(self._labels)), self._labels))

        mod = types.ModuleType("sklearn.cross_validation")
        mod.StratifiedKFold = StratifiedKFold
        sys.modules["sklearn.cross_validation"] = mod
    ns = types.SimpleNamespace()
    base = "BNMTF_ARD.code.cross_validation."
    ns.mask = importlib.import_module(base + "mask")
    ns.MatrixCrossValidation = importlib.import_module(base + "matrix_cross_validation").MatrixCrossValidation
    ns.MatrixNestedCrossValidation = importlib.import_module(base + "nested_matrix_cross_validation").MatrixNestedCrossValidation
    ns.MatrixSingleCrossValidation = importlib.import_module(base + "matrix_single_cross_validation").MatrixSingleCrossValidation
    return ns
