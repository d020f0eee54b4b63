"""Data ingest against golden outputs of the reference's loader (oracle/gen_golden_data.py)."""
import os

import numpy as np

from bnmtf_ard_b200.data import load_data

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REF_CCLE = "/root/reference/data/drug_sensitivity/CCLE/processed_all/"


def test_load_data_create_mask_golden():
    g = np.load(os.path.join(GOLD, "data_golden.npz"))
    path = os.path.join(GOLD, "data_fixture.txt")
    R, M = load_data.load_data_create_mask(path)
    np.testing.assert_array_equal(R, g["fix_R"])
    np.testing.assert_array_equal(M, g["fix_M"])
    assert not np.isnan(R).any() and set(np.unique(M)) <= {0.0, 1.0}
    R, M = load_data.load_gdsc_ic50(path)
    np.testing.assert_array_equal(R, g["gdsc_R"])
    np.testing.assert_array_equal(M, g["gdsc_M"])
    assert R.shape[0] == int((g["fix_M"].sum(axis=1) >= 3).sum()) < g["fix_R"].shape[0]
    for fn in (load_data.load_ctrp_ec50, load_data.load_ccle_ic50, load_data.load_ccle_ec50):
        R2, M2 = fn(path)
        np.testing.assert_array_equal(R2, g["fix_R"])
        np.testing.assert_array_equal(M2, g["fix_M"])


def test_ccle_checksums_when_reference_present():
    """Build container only: the reference's CCLE files through our loader give the reference's sums."""
    if not os.path.isdir(REF_CCLE):
        return
    g = np.load(os.path.join(GOLD, "data_golden.npz"))
    for name, fn in (("ic50", load_data.load_ccle_ic50), ("ec50", load_data.load_ccle_ec50)):
        R, M = fn(REF_CCLE + name + ".txt")
        got = np.array([R.shape[0], R.shape[1], M.sum(), R.sum(), (R * R).sum()])
        np.testing.assert_allclose(got, g["ccle_%s_summary" % name], rtol=1e-13)
