#!/usr/bin/env python
"""Per-sweep time of the two-factor Gibbs sweep on a few shapes, per-row-chain kernel (kernels_v5.cuh) against the
lockstep cluster kernel (kernels_v2.cuh): python tools/v5_shapes.py  (run once with BNMTF_V5=0 and once without)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
sys.path.insert(0, os.path.join(ROOT, "tools"))
from bench_configs import synth
from bnmtf_ard_b200 import _lib
from bnmtf_ard_b200._engine import NMFEngine

dev = torch.device("cuda", 0)
out = {}
for (I, J, K, dens) in [(10000, 8000, 20, 0.3), (6000, 3000, 10, 0.25), (4000, 16000, 30, 0.1), (8000, 8000, 50, 0.5)]:
    R, M = synth(I, J, K, None, dens, dev)
    e = NMFEngine.from_device_blocks(I, J, K, R.data_ptr(), M.data_ptr(), R.data_ptr(), M.data_ptr(), device=0)
    del R, M
    e.set_hyper(True, 1., 1., 1., 1.)
    e.set_scalars(1.0, np.ones(K)); e.init_factors(_lib.GIBBS, True, 1)
    e.run(_lib.GIBBS, 3, seed=1, traces=False)
    e.time_sweeps(True)
    e.run(_lib.GIBBS, 5, seed=2, traces=False)
    t = e.sweep_times(); info = e.layout_info()
    key = "%dx%d K=%d rho=%.2f" % (I, J, K, dens)
    out[key] = {"ms_U": t["ms_U"] / t["n_U"], "ms_V": t["ms_V"] / t["n_V"], "kernel_U": int(info["v2_U"]), "kernel_V": int(info["v2_V"]),
                "npt_U": int(info["v2_npt_U"]), "npt_V": int(info["v2_npt_V"]), "rs_U": int(info["v2_rows_per_set_U"])}
    print(key, json.dumps(out[key]))
    e.close()
