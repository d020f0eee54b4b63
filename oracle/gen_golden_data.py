"""
TEST INFRASTRUCTURE ONLY.  Golden vectors for the data-ingest functions: runs the REAL reference
loader (data/drug_sensitivity/load_data.py, imported from /root/reference) on a small
tab-separated fixture with nan entries and on the reference's own CCLE files, and stores inputs
and outputs in tests/golden/data_golden.npz (+ the fixture text file).  Run in the build container:
    python oracle/gen_golden_data.py
"""
import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "..", "tests", "golden")
REF = "/root/reference/data/drug_sensitivity/load_data.py"


def main():
    spec = importlib.util.spec_from_file_location("ref_load_data", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    rng = np.random.default_rng(77)
    X = np.round(rng.normal(0, 3, (9, 6)), 4)
    X[rng.random(X.shape) < 0.35] = np.nan
    X[3, :] = np.nan
    X[3, 1] = 1.5                       # 1 observed entry: dropped by the GDSC filter
    X[5, :4] = np.nan                   # 2 observed entries: dropped as well
    path = os.path.join(GOLD, "data_fixture.txt")
    np.savetxt(path, X, delimiter="\t", fmt="%.4f")
    out = {}
    out["fix_R"], out["fix_M"] = ref.load_data_create_mask(path)
    out["gdsc_R"], out["gdsc_M"] = ref.load_gdsc_ic50(path)
    # the reference's own CCLE matrices (504 x 24): checksums only, the files stay in the reference
    for name, fn in (("ccle_ic50", ref.load_ccle_ic50), ("ccle_ec50", ref.load_ccle_ec50)):
        R, M = fn()
        out[name + "_summary"] = np.array([R.shape[0], R.shape[1], M.sum(), R.sum(), (R * R).sum()])
    np.savez_compressed(os.path.join(GOLD, "data_golden.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
