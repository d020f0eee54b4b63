"""The reference's OWN known-answer test-suite (tests/code/test_*.py, 122 tests: SURVEY.md section 4, "the parity pins")
executed against the drop-in classes of this package.

``oracle/make_ref.py`` keeps a runnable copy of the reference -- tests included -- under ``oracle/_ref`` (git-ignored,
shipped to the GPU box with the snapshot).  Here every ``BNMTF_ARD.code.models...`` import of those test files is
redirected to ``bnmtf_ard_b200...`` (``sys.modules`` aliases installed only while a test file is being imported), and
each of their test functions becomes one parametrised case below, so the pytest report lists them one by one.

Expected failures are itemised in ``EXPECTED`` with the reason; the only accepted class is a literal ``==`` (or 1e-15)
comparison that a different -- but fixed -- float64 summation order cannot meet.  A listed test that passes, or an
unlisted one that fails, fails this file.  A per-test report goes to ``gpurun_out/reference_suite.json`` when that
directory exists.
"""
import ast
import importlib
import json
import os
import sys
import types

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TESTS = os.path.join(ROOT, "oracle", "_ref", "BNMTF_ARD", "tests", "code")
FILES = ["test_bnmf_gibbs.py", "test_bnmf_vb.py", "test_nmf_icm.py", "test_nmf_np.py", "test_bnmtf_gibbs.py",
         "test_bnmtf_vb.py", "test_nmtf_icm.py", "test_nmtf_np.py", "distributions/test_gamma.py",
         "distributions/test_truncated_normal.py", "distributions/test_truncated_normal_vector.py"]

# file::test -> why the literal comparison of the reference cannot be met by a different float64 summation order
EXPECTED = {
    # measured on the B200 with tools/refsuite_diffs.py (gpurun_out/r02_refsuite_diffs.log -> profiles/r02_reference_suite.md)
    "test_bnmf_vb.py::test_exp_square_diff": "`== 172.66666666666666`: the column-sum form gives 172.66666666666669 (1 ulp, rel 1.7e-16)",
    "test_bnmf_vb.py::test_update_tau": "`beta_s == betatau + 172.66666666666666/2.`: same sum as test_exp_square_diff, 1 ulp",
    "test_bnmf_vb.py::test_update_U": "`mu_U[i,k] == ...` literal: residual form + FMA differ from numpy's expression by 1 ulp (rel 1.9e-16); tau_U equal",
    "test_bnmtf_vb.py::test_update_F": "`abs(mu_F - muFik) < 1e-17` on values of order 1: below one ulp of float64, only the identical operation order passes",
    "test_bnmtf_vb.py::test_update_G": "`abs(mu_G - muGjl) < 1e-17`: as test_update_F",
    "test_nmtf_np.py::test_compute_I_div": "`I_div == expected` literal: 2344.289361880598 against 2344.2893618805974 (1 ulp, rel 1.9e-16, fixed-order tree sum)",
}

MODEL_MODULES = ["bnmf_gibbs", "bnmf_vb", "nmf_icm", "nmf_np", "bnmtf_gibbs", "bnmtf_vb", "nmtf_icm", "nmtf_np"]
DIST_MODULES = ["exponential", "gamma", "normal", "truncated_normal", "truncated_normal_vector"]


def _cases():
    out = []
    for f in FILES:
        path = os.path.join(REF_TESTS, f)
        if not os.path.exists(path):
            continue
        with open(path) as fh:
            tree = ast.parse(fh.read())
        out += [(f, n.name) for n in tree.body if isinstance(n, ast.FunctionDef) and n.name.startswith("test_")]
    return out


CASES = _cases()
_loaded = {}
_report = {}


def _load(fname):
    """Import one reference test file with BNMTF_ARD.code.models.* aliased onto this package."""
    if fname in _loaded:
        return _loaded[fname]
    import bnmtf_ard_b200 as pkg
    saved = {k: v for k, v in sys.modules.items() if k == "BNMTF_ARD" or k.startswith("BNMTF_ARD.")}
    for k in saved:
        del sys.modules[k]
    try:
        top, code = types.ModuleType("BNMTF_ARD"), types.ModuleType("BNMTF_ARD.code")
        top.__path__, code.__path__ = [], []
        top.code, code.models = code, pkg
        sys.modules.update({"BNMTF_ARD": top, "BNMTF_ARD.code": code, "BNMTF_ARD.code.models": pkg})
        for name in MODEL_MODULES:
            sys.modules["BNMTF_ARD.code.models." + name] = importlib.import_module("bnmtf_ard_b200." + name)
        sys.modules["BNMTF_ARD.code.models.distributions"] = importlib.import_module("bnmtf_ard_b200.distributions")
        for name in DIST_MODULES:
            sys.modules["BNMTF_ARD.code.models.distributions." + name] = importlib.import_module("bnmtf_ard_b200.distributions." + name)
        path = os.path.join(REF_TESTS, fname)
        mod = types.ModuleType("refsuite_" + fname.replace("/", "_")[:-3])
        mod.__file__ = path
        with open(path) as fh:
            exec(compile(fh.read(), path, "exec"), mod.__dict__)
    finally:
        for k in [k for k in sys.modules if k == "BNMTF_ARD" or k.startswith("BNMTF_ARD.")]:
            del sys.modules[k]
        sys.modules.update(saved)
    _loaded[fname] = mod
    return mod


def test_the_reference_suite_is_present():
    assert os.path.isdir(REF_TESTS), "oracle/_ref is missing: run `python oracle/make_ref.py` (build() does) before shipping to the GPU box"
    assert len(CASES) >= 120, "expected the reference's 122 tests, found %d" % len(CASES)


@pytest.mark.parametrize("fname,test", CASES, ids=["%s::%s" % c for c in CASES])
def test_reference_suite(fname, test):
    key = "%s::%s" % (fname, test)
    mod = _load(fname)
    fn = getattr(mod, test)
    try:
        with _quiet():
            fn()
    except BaseException as e:   # noqa: BLE001  (pytest.raises inside the reference tests re-raises AssertionError)
        _report[key] = "fail: %s: %s" % (type(e).__name__, str(e)[:300])
        _dump()
        if key in EXPECTED:
            pytest.xfail(EXPECTED[key])
        raise
    _report[key] = "pass"
    _dump()
    assert key not in EXPECTED, "listed as an expected failure but passes: remove it from EXPECTED"


class _quiet(object):
    """The reference classes print one line per iteration; keep the pytest output readable."""

    def __enter__(self):
        import io
        self._out = sys.stdout
        sys.stdout = io.StringIO()

    def __exit__(self, *a):
        sys.stdout = self._out
        return False


def _dump():
    out_dir = os.path.join(ROOT, "gpurun_out")
    if not os.path.isdir(out_dir):
        return
    n_pass = sum(1 for v in _report.values() if v == "pass")
    with open(os.path.join(out_dir, "reference_suite.json"), "w") as fh:
        json.dump({"total_reference_tests": len(CASES), "run": len(_report), "passed": n_pass,
                   "failed": {k: v for k, v in _report.items() if v != "pass"}}, fh, indent=1, sort_keys=True)
