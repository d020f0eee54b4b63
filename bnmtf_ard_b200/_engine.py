"""numpy <-> C-ABI adapter for one NMF model handle (include/bnmtf_b200.h)."""
import ctypes as C

import os

import numpy as np

from . import _dist, _lib
from ._lib import (F_MU, F_TAU, F_VALUE, F_VAR, NSTATS, SIDE_U, SIDE_V, check, dptr, f64)


def numpy_copy(v):
    return np.array(v, dtype=np.float64)


class NMFEngine(object):
    """Owns a bnmtf_nmf_t.  R, M are the full I x J matrices (every rank passes the
    same ones, like the reference constructors); this rank packs its row shard for
    the U sweep and its column shard for the V sweep."""

    def __init__(self, R, M, K, device=None, distributed=True, L=None, layout_floor=None):
        lib = _lib.require_device()
        self.lib = lib
        self.layout_floor = layout_floor
        self.I, self.J = R.shape
        self.K = int(K)
        self.L = None if L is None else int(L)
        self.device = _lib.current_device() if device is None else int(device)
        self.rank, self.world = _dist.info() if distributed else (0, 1)
        r0, r1, rs = _dist.shard_bounds(self.I, self.rank, self.world)
        c0, c1, cs = _dist.shard_bounds(self.J, self.rank, self.world)
        R = f64(R)
        M8 = np.ascontiguousarray(np.asarray(M) != 0, dtype=np.uint8)
        h = C.c_void_p()
        if self.L is None:
            check(lib.bnmtf_nmf_create(C.byref(h), self.device, self.I, self.J, self.K, dptr(R),
                                       M8.ctypes.data_as(C.POINTER(C.c_uint8)), r0, r1, rs, c0, c1, cs))
        else:
            check(lib.bnmtf_nmtf_create(C.byref(h), self.device, self.I, self.J, self.K, self.L, dptr(R),
                                        M8.ctypes.data_as(C.POINTER(C.c_uint8)), r0, r1, rs, c0, c1, cs))
        self.h = h
        self._finish_setup()

    @classmethod
    def from_device_blocks(cls, I, J, K, Rrow_ptr, Mrow_ptr, Rcol_ptr, Mcol_ptr, device=None, L=None):
        """Build from dense blocks already on the device (raw device addresses): rows
        [r0,r1) x J and I x columns [c0,c1) of this rank's shards (see shard_bounds)."""
        self = cls.__new__(cls)
        lib = _lib.require_device()
        self.lib = lib
        self.layout_floor = None
        self.I, self.J, self.K = int(I), int(J), int(K)
        self.L = None if L is None else int(L)
        self.device = _lib.current_device() if device is None else int(device)
        self.rank, self.world = _dist.info()
        r0, r1, rs = _dist.shard_bounds(self.I, self.rank, self.world)
        c0, c1, cs = _dist.shard_bounds(self.J, self.rank, self.world)
        h = C.c_void_p()
        ptrs = (C.c_void_p(Rrow_ptr), C.c_void_p(Mrow_ptr), C.c_void_p(Rcol_ptr), C.c_void_p(Mcol_ptr))
        if self.L is None:
            check(lib.bnmtf_nmf_create_dev(C.byref(h), self.device, self.I, self.J, self.K, *ptrs, r0, r1, rs, c0, c1, cs))
        else:
            check(lib.bnmtf_nmtf_create_dev(C.byref(h), self.device, self.I, self.J, self.K, self.L, *ptrs,
                                            r0, r1, rs, c0, c1, cs))
        self.h = h
        self._finish_setup()
        return self

    def _finish_setup(self):
        lib = self.lib
        # global layout statistics (MAX over ranks) -> sharding-independent packing
        st = np.zeros(4, dtype=np.int64)
        check(lib.bnmtf_nmf_layout_stats(self.h, st.ctypes.data_as(C.POINTER(C.c_int64))))
        if self.world > 1:
            st = _dist.allreduce_max(st)
        if self.layout_floor is not None:   # models of one batch share one row-team layout (bnmtf_batch_run)
            st = np.maximum(st, np.asarray(self.layout_floor, dtype=np.int64))
        check(lib.bnmtf_nmf_pack(self.h, st.ctypes.data_as(C.POINTER(C.c_int64))))
        out = np.zeros(3)
        check(lib.bnmtf_nmf_data_stats(self.h, dptr(out)))
        reduce = _dist.allreduce_sum if self.world > 1 else (lambda v: numpy_copy(v))
        tot = reduce(out[:2])
        self.size_Omega = float(tot[0])
        self.mean_R = float(tot[1] / tot[0]) if tot[0] > 0 else 0.0
        check(lib.bnmtf_nmf_set_data_stats(self.h, self.size_Omega, self.mean_R))
        check(lib.bnmtf_nmf_data_stats(self.h, dptr(out)))
        self.ss_total = float(reduce(out[2:3])[0])
        if self.world > 1:
            ident = np.zeros(128, dtype=np.uint8)
            if self.rank == 0:
                check(lib.bnmtf_nccl_unique_id(ident.ctypes.data_as(C.POINTER(C.c_uint8))))
            ident = np.frombuffer(_dist.broadcast_bytes(ident.tobytes(), src=0), dtype=np.uint8).copy()
            check(lib.bnmtf_nmf_set_comm(self.h, self.rank, self.world, ident.ctypes.data_as(C.POINTER(C.c_uint8))))
            # two-factor models on NCCL ranks of one node: the sweeps write their rows straight into the other ranks'
            # arrays (CUDA IPC) and a device-side barrier replaces the all-gathers; BNMTF_PEER=0 keeps the NCCL calls
            td = _dist._pg()
            # (measured at 20000 x 20000, K = 50: 537 against 526 it/s on 8 GPUs, 141 against 146 on 2 -- the three device
            # barriers per iteration cost more than two small all-gathers there -- so the default is 4 ranks and more;
            # BNMTF_PEER=1 / 0 forces it on / off)
            want = os.environ.get("BNMTF_PEER", "")
            peer = (want == "1") or (want != "0" and self.world >= 4)
            if self.L is None and td is not None and td.get_backend() == "nccl" and peer:
                nb = 448
                blob = np.zeros(nb, dtype=np.uint8)
                check(lib.bnmtf_nmf_peer_export(self.h, blob.ctypes.data_as(C.POINTER(C.c_uint8)), nb))
                blobs = np.frombuffer(_dist.allgather_bytes(blob.tobytes()), dtype=np.uint8).copy()
                check(lib.bnmtf_nmf_peer_import(self.h, blobs.ctypes.data_as(C.POINTER(C.c_uint8)), nb))

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.lib.bnmtf_nmf_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    # -- state ---------------------------------------------------------------
    def n_of(self, side):
        return self.I if side == SIDE_U else self.J

    def k_of(self, side):
        return self.K if (side == SIDE_U or self.L is None) else self.L

    def set_hyper(self, ARD, alphatau, betatau, alpha0=1.0, beta0=1.0, lambdaU=None, lambdaV=None):
        lu = None if lambdaU is None else f64(lambdaU, (self.I, self.K))
        lv = None if lambdaV is None else f64(lambdaV, (self.J, self.K))
        check(self.lib.bnmtf_nmf_set_hyper(self.h, int(bool(ARD)), float(alphatau), float(betatau), float(alpha0),
                                           float(beta0), dptr(lu), dptr(lv)))

    def set_factor(self, side, field, arr):
        a = f64(arr, (self.n_of(side), self.k_of(side)))
        check(self.lib.bnmtf_nmf_set_factor(self.h, side, field, dptr(a)))
        if self.world > 1:
            self._uploaded = getattr(self, "_uploaded", 0.0) + float(a.sum()) * (1 + side + 2 * field)

    def _check_replicated_state(self):
        """One process per GPU: every rank uploads its own host copy of the state and the shards exchanged during the
        run are conditioned on it, so the copies must be identical -- a cheap fingerprint of what was uploaded since
        the last run is compared over the ranks before the device loop starts."""
        if self.world > 1:
            fp, self._uploaded = getattr(self, "_uploaded", 0.0), 0.0
            if not _dist.same_on_all_ranks([fp]):
                raise RuntimeError("multi-GPU run: the ranks hold different initial states (factor fingerprint %r on rank %d); "
                                   "initialise with the same values on every rank -- initialise() of the model classes does, "
                                   "its seed is broadcast from rank 0" % (fp, self.rank))

    def get_factor(self, side, field):
        out = np.empty((self.n_of(side), self.k_of(side)))
        check(self.lib.bnmtf_nmf_get_factor(self.h, side, field, dptr(out)))
        return out

    def set_scalars(self, tau, lambdak=None):
        lk = None if lambdak is None else f64(lambdak, (self.K,))
        check(self.lib.bnmtf_nmf_set_scalars(self.h, float(tau), dptr(lk)))
        if self.world > 1:
            self._uploaded = getattr(self, "_uploaded", 0.0) + 7.0 * float(tau) + (0.0 if lk is None else 11.0 * float(lk.sum()))

    def get_scalars(self, want_lambdak=True):
        tau = C.c_double()
        lk = np.zeros(self.K) if want_lambdak else None
        check(self.lib.bnmtf_nmf_get_scalars(self.h, C.byref(tau), dptr(lk)))
        return tau.value, lk

    def set_vb_scalars(self, exp_tau, exp_logtau, alpha_s, beta_s, alphak_s=None, betak_s=None, exp_lambdak=None,
                       exp_loglambdak=None):
        vecs = [None if v is None else f64(v, (self.K,)) for v in (alphak_s, betak_s, exp_lambdak, exp_loglambdak)]
        check(self.lib.bnmtf_nmf_set_vb_scalars(self.h, float(exp_tau), float(exp_logtau), float(alpha_s),
                                                float(beta_s), *[dptr(v) for v in vecs]))

    def get_vb_scalars(self):
        sc, ls = np.zeros(4), np.zeros((4, self.K))
        check(self.lib.bnmtf_nmf_get_vb_scalars(self.h, dptr(sc), dptr(ls)))
        return sc, ls

    def init_factors(self, method, random, seed):
        check(self.lib.bnmtf_nmf_init_factors(self.h, method, int(bool(random)), C.c_uint64(seed)))

    # -- compute ---------------------------------------------------------------
    def column_params(self, method, side, k):
        n = self.n_of(side)
        t, m = np.empty(n), np.empty(n)
        check(self.lib.bnmtf_nmf_column_params(self.h, method, side, int(k), dptr(t), dptr(m)))
        return t, m

    def stats(self, method):
        out = np.zeros(NSTATS)
        check(self.lib.bnmtf_nmf_stats(self.h, method, dptr(out)))
        return out

    def run(self, method, iterations, seed=0, traces=True, want_lambdak=True):
        it = int(iterations)
        self._check_replicated_state()
        res = {
            "all_U": np.zeros((it, self.I, self.K)) if traces else None,
            "all_V": np.zeros((it, self.J, self.K)) if traces else None,
            "all_lambdak": np.zeros((it, self.K)) if want_lambdak else None,
            "stats": np.zeros((it, NSTATS)),
            "seconds": np.zeros(it),
        }
        check(self.lib.bnmtf_nmf_run(self.h, method, it, C.c_uint64(seed), dptr(res["all_U"]), dptr(res["all_V"]),
                                     dptr(res["all_lambdak"]), dptr(res["stats"]), dptr(res["seconds"])))
        return res

    # -- on-device posterior summaries / held-out prediction --------------------
    def set_summary(self, burn_in, thinning):
        check(self.lib.bnmtf_nmf_set_summary(self.h, int(burn_in), int(thinning)))

    def summary_count(self):
        n = C.c_int64(0)
        check(self.lib.bnmtf_nmf_get_summary(self.h, 0, None, C.byref(n)))
        return int(n.value)

    def get_summary(self, field):
        L = self.K if self.L is None else self.L
        shape = {_lib.SUM_FACTOR_U: (self.I, self.K), _lib.SUM_FACTOR_V: (self.J, L), _lib.SUM_S: (self.K, L),
                 _lib.SUM_LAMBDA_U: (self.K,), _lib.SUM_LAMBDA_V: (L,), _lib.SUM_TAU: (1,)}[field]
        out = np.empty(shape)
        check(self.lib.bnmtf_nmf_get_summary(self.h, int(field), dptr(out), None))
        return out

    def predict_entries(self, source, rows, cols, vals):
        """compute_MSE / compute_R2 / compute_Rp (bnmf_gibbs.py:252-270) over the listed entries."""
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        cols = np.ascontiguousarray(cols, dtype=np.int32)
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        out = np.zeros(8)
        i32 = C.POINTER(C.c_int32)
        check(self.lib.bnmtf_nmf_predict(self.h, int(source), rows.size, rows.ctypes.data_as(i32),
                                         cols.ctypes.data_as(i32), dptr(vals), dptr(out)))
        return {"n": out[0], "MSE": float(out[1]), "R^2": float(out[2]), "Rp": float(out[3]), "SS_res": float(out[4]),
                "SS_total": float(out[5])}

    def launch_count(self):
        return int(self.lib.bnmtf_nmf_launch_count(self.h))

    def layout_info(self):
        out = np.zeros(16, dtype=np.int64)
        check(self.lib.bnmtf_nmf_layout_info(self.h, out.ctypes.data_as(C.POINTER(C.c_int64))))
        names = ["tpr_U", "npt_U", "tpr_V", "npt_V", "slots_U", "slots_V", "nnz_U", "nnz_V",
                 "v2_U", "v2_npt_U", "v2_rows_per_set_U", "v2_slots_U", "v2_V", "v2_npt_V", "v2_rows_per_set_V", "v2_slots_V"]
        return dict(zip(names, out.tolist()))

    def time_sweeps(self, on):
        """Measurement passes only: bracket every sweep launch by CUDA events (serialises the run loop)."""
        check(self.lib.bnmtf_nmf_time_sweeps(self.h, int(bool(on))))

    def sweep_times(self):
        out = np.zeros(4)
        check(self.lib.bnmtf_nmf_sweep_times(self.h, dptr(out)))
        return {"ms_U": out[0], "ms_V": out[1], "n_U": int(out[2]), "n_V": int(out[3])}

    # -- metrics from the per-iteration sufficient statistics ------------------
    def performances(self, stats_row):
        """MSE, R^2, Rp on Omega (bnmf_gibbs.py:252-270) from RSS, sum E and
        sum (R-mean)E; algebraically identical to the reference's centred sums."""
        rss, se, srce = stats_row[_lib.ST_RSS], stats_row[_lib.ST_SUM_E], stats_row[_lib.ST_SUM_RCE]
        n, sst = self.size_Omega, self.ss_total
        mse = rss / n
        r2 = 1.0 - rss / sst if sst != 0.0 else np.inf
        cov = sst - srce
        var_pred = sst - 2.0 * srce + (rss - se * se / n)
        with np.errstate(all="ignore"):
            rp = cov / (np.sqrt(sst) * np.sqrt(var_pred))
        return {"MSE": float(mse), "R^2": float(r2), "Rp": float(rp)}


    def performances_all(self, stats):
        """performances() of every row of a (iterations, NSTATS) block; element for element the same
        float64 expressions, evaluated by numpy over the iteration axis."""
        stats = np.asarray(stats, dtype=np.float64).reshape(-1, NSTATS)
        rss, se, srce = stats[:, _lib.ST_RSS], stats[:, _lib.ST_SUM_E], stats[:, _lib.ST_SUM_RCE]
        n, sst = self.size_Omega, self.ss_total
        mse = rss / n
        r2 = 1.0 - rss / sst if sst != 0.0 else np.full(len(rss), np.inf)
        cov = sst - srce
        var_pred = sst - 2.0 * srce + (rss - se * se / n)
        with np.errstate(all="ignore"):
            rp = cov / (np.sqrt(sst) * np.sqrt(var_pred))
        return {"MSE": [float(v) for v in mse], "R^2": [float(v) for v in r2], "Rp": [float(v) for v in rp]}


class NMTFEngine(NMFEngine):
    """Handle for the tri-factorisation R ~ F S G^T: side U is F (I x K), side V is G (J x L)."""

    def __init__(self, R, M, K, L, device=None, distributed=True):
        NMFEngine.__init__(self, R, M, K, device=device, distributed=distributed, L=L)

    def set_hyper(self, ARD, alphatau, betatau, alpha0=1.0, beta0=1.0, lambdaF=None, lambdaG=None, lambdaS=None):
        lf = None if lambdaF is None else f64(lambdaF, (self.I, self.K))
        lg = None if lambdaG is None else f64(lambdaG, (self.J, self.L))
        ls = f64(lambdaS, (self.K, self.L))
        check(self.lib.bnmtf_nmtf_set_hyper(self.h, int(bool(ARD)), float(alphatau), float(betatau), float(alpha0),
                                            float(beta0), dptr(lf), dptr(lg), dptr(ls)))

    def set_S(self, field, arr):
        a = f64(arr, (self.K, self.L))
        check(self.lib.bnmtf_nmtf_set_S(self.h, field, dptr(a)))

    def get_S(self, field):
        out = np.empty((self.K, self.L))
        check(self.lib.bnmtf_nmtf_get_S(self.h, field, dptr(out)))
        return out

    def set_scalars(self, tau, lambdaFk=None, lambdaGl=None):
        lf = None if lambdaFk is None else f64(lambdaFk, (self.K,))
        lg = None if lambdaGl is None else f64(lambdaGl, (self.L,))
        check(self.lib.bnmtf_nmtf_set_scalars(self.h, float(tau), dptr(lf), dptr(lg)))

    def init(self, method, random_FG, random_S, seed_FG, seed_S):
        check(self.lib.bnmtf_nmtf_init(self.h, method, int(bool(random_FG)), int(bool(random_S)),
                                       C.c_uint64(seed_FG), C.c_uint64(seed_S)))

    def column_params(self, method, which, k=0, l=0):
        n = (self.I, self.J, 1)[which]
        t, m = np.empty(n), np.empty(n)
        check(self.lib.bnmtf_nmtf_column_params(self.h, method, which, int(k), int(l), dptr(t), dptr(m)))
        return t, m

    def stats(self, method):
        out = np.zeros(NSTATS)
        check(self.lib.bnmtf_nmtf_stats(self.h, method, dptr(out)))
        return out

    def run(self, method, iterations, seed=0, traces=True):
        it = int(iterations)
        self._check_replicated_state()
        res = {
            "all_F": np.zeros((it, self.I, self.K)) if traces else None,
            "all_S": np.zeros((it, self.K, self.L)) if traces else None,
            "all_G": np.zeros((it, self.J, self.L)) if traces else None,
            "all_lambdaFk": np.zeros((it, self.K)),
            "all_lambdaGl": np.zeros((it, self.L)),
            "stats": np.zeros((it, NSTATS)),
            "seconds": np.zeros(it),
        }
        check(self.lib.bnmtf_nmtf_run(self.h, method, it, C.c_uint64(seed), dptr(res["all_F"]), dptr(res["all_S"]),
                                      dptr(res["all_G"]), dptr(res["all_lambdaFk"]), dptr(res["all_lambdaGl"]),
                                      dptr(res["stats"]), dptr(res["seconds"])))
        return res

    # -- VB (bnmtf_vb) -------------------------------------------------------------
    def set_vb_scalars(self, exp_tau, exp_logtau, alpha_s, beta_s, stateF=None, stateG=None):
        sf = None if stateF is None else f64(stateF, (4, self.K))
        sg = None if stateG is None else f64(stateG, (4, self.L))
        check(self.lib.bnmtf_nmtf_set_vb_scalars(self.h, float(exp_tau), float(exp_logtau), float(alpha_s), float(beta_s),
                                                 dptr(sf), dptr(sg)))

    def get_vb_scalars(self):
        sc, sf, sg = np.zeros(4), np.zeros((4, self.K)), np.zeros((4, self.L))
        check(self.lib.bnmtf_nmtf_get_vb_scalars(self.h, dptr(sc), dptr(sf), dptr(sg)))
        return sc, sf, sg

    def vb_update_params(self, which, k=0, l=0):
        n = (self.I, self.J, 1)[which]
        t, m = np.empty(n), np.empty(n)
        check(self.lib.bnmtf_nmtf_vb_update_params(self.h, which, int(k), int(l), dptr(t), dptr(m)))
        return t, m

    def vb_stats(self):
        out = np.zeros(NSTATS)
        check(self.lib.bnmtf_nmtf_vb_stats(self.h, dptr(out)))
        return out

    def vb_run(self, iterations):
        it = int(iterations)
        res = {"stats": np.zeros((it, NSTATS)), "seconds": np.zeros(it)}
        check(self.lib.bnmtf_nmtf_vb_run(self.h, it, dptr(res["stats"]), dptr(res["seconds"])))
        return res

    # -- NP (nmtf_np) ----------------------------------------------------------------
    def np_update(self, which, k=0, l=0):
        out = np.empty((self.I, self.J, 1)[which])
        check(self.lib.bnmtf_nmtf_np_update(self.h, which, int(k), int(l), dptr(out)))
        return out

    def np_run(self, iterations):
        it = int(iterations)
        res = {"stats": np.zeros((it, NSTATS)), "seconds": np.zeros(it)}
        check(self.lib.bnmtf_nmtf_np_run(self.h, it, dptr(res["stats"]), dptr(res["seconds"])))
        return res
