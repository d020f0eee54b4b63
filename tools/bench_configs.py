#!/usr/bin/env python
"""Timing of the other BASELINE.json configurations (the bench.py line is configs[2]):
  C1  BNMF Gibbs+ARD, toy-shaped 100 x 80, K=10, all observed
  C2  BNMF VB+ARD, GDSC-shaped 707 x 139, 81 % observed, K=20
  C4  BNMTF Gibbs and VB, 10000 x 8000, 30 % observed, K=L=20
Synthetic data of each shape; iterations/s from CUDA events inside the engine.
Usage: python tools/bench_configs.py [--out profiles/rNN_configs.json]
C4 also runs sharded: python -m torch.distributed.run --nproc-per-node N ... tools/bench_configs.py --only C4
(one rank per GPU: rows of R for the F sweep, columns for the G sweep, as bench.py does for configs[2])."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def synth(I, J, K, L, density, device, seed=7):
    import torch
    g = torch.Generator(device=device); g.manual_seed(seed)
    ex = lambda *s: -torch.log(1.0 - torch.rand(*s, generator=g, device=device, dtype=torch.float64))  # noqa: E731
    if L is None:
        R = ex(I, K) @ ex(J, K).T
    else:
        R = ex(I, K) @ ex(K, L) @ ex(J, L).T
    R += torch.randn(I, J, generator=g, device=device, dtype=torch.float64)
    M = (torch.rand(I, J, generator=g, device=device) < density).to(torch.uint8)
    M[torch.arange(I, device=device), torch.randint(0, J, (I,), generator=g, device=device)] = 1
    M[torch.randint(0, I, (J,), generator=g, device=device), torch.arange(J, device=device)] = 1
    return R.contiguous(), M.contiguous()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--only", default="", help="substring filter on the configuration name (C1, C2, C4)")
    ap.add_argument("--its", type=int, default=0, help="override the iteration counts (profiling)")
    args = ap.parse_args()
    import torch
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        args.only = args.only or "C4"
        assert "C4" in args.only, "only the C4 configurations run sharded"
    from bnmtf_ard_b200 import _dist, _lib
    from bnmtf_ard_b200._engine import NMFEngine, NMTFEngine
    from bnmtf_ard_b200._lib import F_MU, F_TAU, F_VALUE, F_VAR, SIDE_U, SIDE_V
    dev = torch.device("cuda", local)
    out = {}

    def nmf_engine(I, J, K, dens):
        R, M = synth(I, J, K, None, dens, dev)
        e = NMFEngine.from_device_blocks(I, J, K, R.data_ptr(), M.data_ptr(), R.data_ptr(), M.data_ptr(), device=0)
        e.set_hyper(True, 1., 1., 1., 1.)
        return e

    def record(name, seconds, its, extra=None):
        s = float(seconds[-1] - seconds[max(0, len(seconds) - its - 1)]) if len(seconds) > its else float(seconds[-1])
        if world > 1:      # the slowest rank
            t = torch.tensor([s], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            s = float(t.item())
        d = {"iterations_per_s": its / s, "ms_per_iteration": 1e3 * s / its, "timed_iterations": its, "n_gpus": world}
        d.update(extra or {})
        out[name] = d
        if rank == 0:
            print(name, json.dumps(d))

    want = lambda tag: (not args.only) or (tag in args.only)  # noqa: E731
    n = lambda default: args.its if args.its > 0 else default  # noqa: E731
    if want("C1"):   # toy-shaped BNMF Gibbs + ARD
        e = nmf_engine(100, 80, 10, 1.1)
        e.set_scalars(1.0, np.ones(10)); e.init_factors(_lib.GIBBS, True, 1)
        r = e.run(_lib.GIBBS, n(1000) + 100, seed=3, traces=True)
        record("C1 BNMF Gibbs+ARD 100x80 K=10", r["seconds"], n(1000), {"reference_published_it_per_s": 21.1, "train_mse_last": e.performances(r["stats"][-1])["MSE"]})
        e.close()
    if want("C2"):   # GDSC-shaped BNMF VB + ARD
        I, J, K = 707, 139, 20
        e = nmf_engine(I, J, K, 0.81)
        e.set_scalars(0.0, np.ones(K)); e.init_factors(_lib.VB, True, 2)
        e.set_vb_scalars(1.0, 0.0, 1.0, 1.0, np.ones(K), np.ones(K), np.ones(K), np.zeros(K))
        r = e.run(_lib.VB, n(200) + 20, seed=0, traces=False)
        record("C2 BNMF VB+ARD 707x139 K=20", r["seconds"], n(200), {"reference_published_it_per_s": 7.65, "elbo_last": float(r["stats"][-1, _lib.ST_ELBO])})
        e.close()
    if want("C4"):   # BNMTF Gibbs and VB
        I, J, K, L = 10000, 8000, 20, 20
        R, M = synth(I, J, K, L, 0.3, dev)
        r0, r1, _ = _dist.shard_bounds(I, rank, world)
        c0, c1, _ = _dist.shard_bounds(J, rank, world)
        Rrow, Mrow = R[r0:r1].contiguous(), M[r0:r1].contiguous()
        Rcol, Mcol = R[:, c0:c1].contiguous(), M[:, c0:c1].contiguous()
        e = NMTFEngine.from_device_blocks(I, J, K, Rrow.data_ptr(), Mrow.data_ptr(), Rcol.data_ptr(), Mcol.data_ptr(), device=local, L=L)
        del R, M, Rrow, Mrow, Rcol, Mcol
        torch.cuda.empty_cache()
        lamS = 0.1 * np.ones((K, L))
        e.set_hyper(True, 1., 1., 1., 1., None, None, lamS)
        if want("gibbs") or args.only in ("", "C4"):
            e.set_scalars(1.0, np.ones(K), np.ones(L))
            e.init(_lib.GIBBS, True, True, 5, 6)
            r = e.run(_lib.GIBBS, n(10) + 3, seed=4, traces=False)
            record("C4 BNMTF Gibbs+ARD 10000x8000 K=L=20", r["seconds"], n(10), {"layout": e.layout_info(), "train_mse_last": e.performances(r["stats"][-1])["MSE"]})
        if want("vb") or args.only in ("", "C4"):
            e.set_scalars(1.0, np.ones(K), np.ones(L))
            e.init(_lib.VB, True, True, 7, 8)
            e.set_vb_scalars(1.0, 0.0, 1.0, 1.0, np.ones((4, K)), np.ones((4, L)))
            r = e.vb_run(n(10) + 3)
            record("C4 BNMTF VB+ARD 10000x8000 K=L=20", r["seconds"], n(10), {"elbo_last": float(r["stats"][-1, _lib.ST_ELBO]), "train_mse_last": e.performances(r["stats"][-1])["MSE"]})
        if want("np") or args.only in ("", "C4"):     # nmtf_np: F sweep, the K*L update_S in one pass each, G sweep (needs R >= 0)
            e.set_scalars(1.0, np.ones(K), np.ones(L))
            e.init(_lib.GIBBS, True, True, 9, 10)       # positive random factors
            r = e.np_run(n(4) + 2)
            record("C4-shaped NMTF NP 10000x8000 K=L=20", r["seconds"], n(4), {"i_div_last": float(r["stats"][-1, _lib.ST_IDIV]) if hasattr(_lib, "ST_IDIV") else None})
        e.close()
    if want("C5"):   # nested-CV-shaped workload: 10 folds x K in {1..30} of BNMF VB (200 it) on a CTRP-shaped 887 x 545 matrix
        import contextlib, io, time
        from bnmtf_ard_b200 import bnmf_vb
        from bnmtf_ard_b200.cross_validation.parallel_matrix_cross_validation import ParallelMatrixCrossValidation
        rng = np.random.default_rng(5)
        I, J = 887, 545
        Rh = rng.exponential(1, (I, 10)) @ rng.exponential(1, (J, 10)).T + rng.normal(0, 1, (I, J))
        Mh = (rng.random((I, J)) < 0.8).astype(float)
        Mh[np.arange(I), rng.integers(0, J, I)] = 1; Mh[rng.integers(0, I, J), np.arange(J)] = 1

        class Quiet(object):
            def __new__(cls, R, M, **p):
                m = bnmf_vb(R, M, **p); m.verbose = False
                return m
        hp = {'alphatau': 1., 'betatau': 1., 'alpha0': 1., 'beta0': 1.}
        Ks = list(range(1, 31)) if args.its <= 0 else list(range(1, 1 + args.its))
        search = [{'K': k, 'ARD': True, 'hyperparameters': hp} for k in Ks]
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            cv = ParallelMatrixCrossValidation(method=Quiet, R=Rh, M=Mh, K=10, parameter_search=search,
                                               train_config={'iterations': 200, 'init_UV': 'random'}, predict_config={},
                                               file_performance="/tmp/cv_c5.txt", P=10)
            cv.run()
            best, perf = cv.find_best_parameters('MSE', True)
        dt = time.perf_counter() - t0
        fits = 10 * len(Ks)
        out["C5 10-fold CV, K=1..%d, BNMF VB 200 it, 887x545" % Ks[-1]] = {
            "seconds": dt, "fits": fits, "fits_per_s": fits / dt, "best_K": best['K'], "best_mse": perf,
            "reference_estimate_s": "0.16-0.59 s/iteration x 200 it x %d fits (BASELINE.md section 2)" % fits}
        print("C5", json.dumps(out[list(out)[-1]]))
        # the same sweep with all fits in one batch (one launch per update for the 300 models)
        from bnmtf_ard_b200.cross_validation.batched_matrix_cross_validation import BatchedMatrixCrossValidation
        for method_name in (["vb", "icm", "np"] if args.its <= 0 else ["vb"]):
            import bnmtf_ard_b200 as pkg
            cls0 = {"vb": pkg.bnmf_vb, "icm": pkg.nmf_icm, "np": pkg.nmf_np}[method_name]

            class QuietB(object):
                def __new__(cls, R, M, **p):
                    m = cls0(R, M, **p); m.verbose = False
                    return m
            if method_name == "np":
                srch = [{'K': k} for k in Ks]; Rb = np.abs(Rh); pc = {}
                tc = {'iterations': 200, 'init_UV': 'random'}
            elif method_name == "icm":
                srch = search; Rb = Rh; pc = {'burn_in': 180, 'thinning': 2}
                tc = {'iterations': 200, 'init_UV': 'random'}
            else:
                srch = search; Rb = Rh; pc = {}
                tc = {'iterations': 200, 'init_UV': 'random'}
            t0 = time.perf_counter()
            with contextlib.redirect_stdout(io.StringIO()):
                cvb = BatchedMatrixCrossValidation(method=QuietB, R=Rb, M=Mh, K=10, parameter_search=srch, train_config=tc,
                                                   predict_config=pc, file_performance="/tmp/cv_c5b.txt")
                cvb.run()
                bestb, perfb = cvb.find_best_parameters('MSE', True)
            dtb = time.perf_counter() - t0
            key = "C5 batched 10-fold CV, K=1..%d, %s 200 it, 887x545" % (Ks[-1], cls0.__name__)
            out[key] = {"seconds": dtb, "fits": fits, "fits_per_s": fits / dtb, "best_K": bestb['K'], "best_mse": perfb,
                        "phases_s": cvb.timing, "launches": cvb.launches}
            print("C5b", json.dumps(out[key]))
    if args.out and rank == 0:
        with open(args.out, "w") as f:
            json.dump(out, f, indent=1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
