"""Multi-GPU parity check (run under torchrun on N GPUs of one box):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
         --master-port 29511 tests/multigpu_check.py

Every rank builds the sharded engine; rank 0 also builds a single-GPU engine on
the same data.  Because the Philox streams are keyed by global row/column
indices and every reduction runs over the gathered global arrays in a fixed
order, the sharded chain must be BITWISE identical to the single-GPU chain
(Gibbs, ICM, VB, NP)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from bnmtf_ard_b200 import _lib
    from bnmtf_ard_b200._engine import NMFEngine
    from bnmtf_ard_b200._lib import F_MU, F_TAU, F_VALUE, F_VAR, SIDE_U, SIDE_V

    ok = True
    # the last shape is a strip of the benchmark configuration (K = 50, rho = 0.2) packed with the benchmark's layout
    # (20 entries per lane, 11 rows per set: the packing floor of the full 20000 x 20000 matrix is handed over)
    for (I, J, K, dens) in [(257, 190, 5, 0.6), (1001, 777, 8, 0.3), (64, 2500, 4, 0.7), (400, 20000, 50, 0.2)]:
        floor = [0, 600, 0, 600] if K == 50 else None
        rng = np.random.default_rng(I + J)
        R = np.abs(rng.exponential(1, (I, K)) @ rng.exponential(1, (J, K)).T + rng.normal(0, 1, (I, J))) + 0.05
        M = (rng.random((I, J)) < dens).astype(float)
        M[np.arange(I), rng.integers(0, J, I)] = 1
        M[rng.integers(0, I, J), np.arange(J)] = 1
        U0, V0 = rng.exponential(1, (I, K)), rng.exponential(1, (J, K))
        engs = [NMFEngine(R, M, K, device=local, layout_floor=floor)]
        if rank == 0:
            engs.append(NMFEngine(R, M, K, device=local, distributed=False, layout_floor=floor))
        if floor is not None:
            info = engs[0].layout_info()
            assert info["v2_U"] >= 1 and info["v2_npt_U"] == 20 and info["v2_rows_per_set_U"] == 11, info
        for method in (_lib.GIBBS, _lib.ICM, _lib.VB, _lib.NP):
            outs = []
            for e in engs:
                e.set_hyper(True, 1., 1., 1., 1.)
                e.set_factor(SIDE_U, F_VALUE, U0); e.set_factor(SIDE_V, F_VALUE, V0)
                if method == _lib.VB:
                    for s, X in ((SIDE_U, U0), (SIDE_V, V0)):
                        e.set_factor(s, F_MU, X); e.set_factor(s, F_TAU, np.ones_like(X)); e.set_factor(s, F_VAR, 0.1 * np.ones_like(X))
                    e.set_vb_scalars(1.2, 0.1, 3., 2.5, np.ones(K), np.ones(K), np.ones(K), np.zeros(K))
                else:
                    e.set_scalars(1.2, np.ones(K))
                res = e.run(method, 4, seed=99, traces=True)
                outs.append((res, e.get_factor(SIDE_U, F_VALUE), e.get_factor(SIDE_V, F_VALUE)))
            if rank == 0:
                a, b = outs
                # the chain (factors, traces, tau, lambda) must be bitwise identical; the
                # R-mean-centred metric sums may differ in the last bits because mean_Omega(R)
                # itself is summed shard by shard
                st_a, st_b = a[0]["stats"], b[0]["stats"]
                same = (np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
                        and np.array_equal(st_a[:, 3], st_b[:, 3]) and np.array_equal(st_a[:, 0], st_b[:, 0])
                        and np.allclose(st_a, st_b, rtol=1e-11, atol=1e-9, equal_nan=True)
                        and np.array_equal(a[0]["all_U"], b[0]["all_U"]) and np.array_equal(a[0]["all_V"], b[0]["all_V"])
                        and np.array_equal(a[0]["all_lambdak"], b[0]["all_lambdak"]))
                print("shape %s method %d world %d: %s (max |dU| %.3e)" % ((I, J, K), method, world, "BITWISE EQUAL" if same else "MISMATCH",
                                                                            np.max(np.abs(a[1] - b[1]))))
                ok = ok and same
        for e in engs:
            e.close()
    # ---- tri-factorisation engines (Gibbs, ICM, VB, NP): F/G sweeps sharded, S statistics gathered
    from bnmtf_ard_b200._engine import NMTFEngine
    # (the last two: F side / G side on the per-row-chain cluster kernel -- projection epilogue, VB covariance term -- sharded)
    for (I, J, K, L, dens) in [(301, 223, 4, 3, 0.6), (90, 2600, 3, 5, 0.5), (2600, 90, 5, 3, 0.5)]:
        rng = np.random.default_rng(I * 7 + J)
        R = np.abs(rng.exponential(1, (I, K)) @ rng.exponential(1, (K, L)) @ rng.exponential(1, (J, L)).T + rng.normal(0, 1, (I, J))) + 0.05
        M = (rng.random((I, J)) < dens).astype(float)
        M[np.arange(I), rng.integers(0, J, I)] = 1
        M[rng.integers(0, I, J), np.arange(J)] = 1
        F0, S0, G0 = rng.exponential(1, (I, K)), rng.exponential(1, (K, L)), rng.exponential(1, (J, L))
        engs = [NMTFEngine(R, M, K, L, device=local)]
        if rank == 0:
            engs.append(NMTFEngine(R, M, K, L, device=local, distributed=False))
        for method in (_lib.GIBBS, _lib.ICM, _lib.VB, _lib.NP):
            outs = []
            for e in engs:
                e.set_hyper(True, 1., 1., 1., 1., None, None, 0.1 * np.ones((K, L)))
                e.set_factor(SIDE_U, F_VALUE, F0); e.set_factor(SIDE_V, F_VALUE, G0); e.set_S(F_VALUE, S0)
                if method == _lib.VB:
                    for s, X in ((SIDE_U, F0), (SIDE_V, G0)):
                        e.set_factor(s, F_MU, X); e.set_factor(s, F_TAU, np.ones_like(X)); e.set_factor(s, F_VAR, 0.1 * np.ones_like(X))
                    e.set_S(F_MU, S0); e.set_S(F_TAU, np.ones_like(S0)); e.set_S(F_VAR, 0.1 * np.ones_like(S0))
                    e.set_vb_scalars(1.2, 0.1, 3., 2.5, np.ones((4, K)), np.ones((4, L)))
                    res = e.vb_run(3)
                elif method == _lib.NP:
                    res = e.np_run(3)
                else:
                    e.set_scalars(1.2, np.ones(K), np.ones(L))
                    res = e.run(method, 3, seed=77, traces=False)
                outs.append((res["stats"], e.get_factor(SIDE_U, F_VALUE), e.get_S(F_VALUE), e.get_factor(SIDE_V, F_VALUE)))
            if rank == 0:
                a, b = outs
                same = (np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])
                        and np.array_equal(a[0][:, 3], b[0][:, 3]) and np.allclose(a[0], b[0], rtol=1e-11, atol=1e-9, equal_nan=True))
                print("NMTF shape %s method %d world %d: %s (max |dF| %.3e, max |dS| %.3e)" % (
                    (I, J, K, L), method, world, "BITWISE EQUAL" if same else "MISMATCH", np.max(np.abs(a[1] - b[1])), np.max(np.abs(a[2] - b[2]))))
                ok = ok and same
        for e in engs:
            e.close()
    # ---- the class API without a common numpy seed (round-1 advice): every rank seeds numpy differently, yet all ranks
    # must hold the same chain -- initialise() / run() take rank 0's Philox seed -- and it must be the single-GPU chain
    from bnmtf_ard_b200 import bnmf_gibbs
    rng = np.random.default_rng(5)
    I, J, K = 300, 260, 6
    R = rng.exponential(1, (I, K)) @ rng.exponential(1, (J, K)).T + rng.normal(0, 1, (I, J))
    M = (rng.random((I, J)) < 0.5).astype(float)
    M[np.arange(I), rng.integers(0, J, I)] = 1
    M[rng.integers(0, I, J), np.arange(J)] = 1
    hyper = {"alphatau": 1., "betatau": 1., "alpha0": 1., "beta0": 1.}
    np.random.seed(1000 + rank)
    m = bnmf_gibbs(R, M, K, True, hyper)
    m.verbose = False
    m.train("random", 12)
    digest = torch.tensor([float(m.U.sum()), float(m.V.sum()), float(m.tau), float(m.all_U[3].sum())], dtype=torch.float64, device="cuda")
    lo, hi = digest.clone(), digest.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    same_ranks = bool(torch.equal(lo, hi))
    if rank == 0:
        from bnmtf_ard_b200 import _dist
        real_info = _dist.info
        _dist.info = lambda: (0, 1)            # a single-GPU model next to the process group
        try:
            np.random.seed(1000)
            s1 = bnmf_gibbs(R, M, K, True, hyper)
            s1.verbose = False
            s1.train("random", 12)
        finally:
            _dist.info = real_info
        same_single = np.array_equal(s1.all_U, m.all_U) and np.array_equal(s1.all_V, m.all_V) and np.array_equal(s1.all_tau, m.all_tau)
        print("class API, numpy seeded per rank, world %d: ranks agree %s, equals the single-GPU chain %s" % (world, same_ranks, same_single))
        ok = ok and same_ranks and same_single
    # ranks that upload different states are refused
    eng = NMFEngine(R, M, K, device=local)
    eng.set_hyper(True, 1., 1., 1., 1.)
    eng.set_factor(SIDE_U, F_VALUE, np.full((I, K), 1.0 + rank)); eng.set_factor(SIDE_V, F_VALUE, np.ones((J, K)))
    eng.set_scalars(1.0, np.ones(K))
    try:
        eng.run(_lib.GIBBS, 1, seed=1, traces=False)
        refused = False
    except RuntimeError as e:
        refused = "different initial states" in str(e)
    if rank == 0:
        print("ranks with different uploaded states are refused: %s" % refused)
        ok = ok and refused
    eng.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTIGPU_CHECK", "PASS" if ok else "FAIL")
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
