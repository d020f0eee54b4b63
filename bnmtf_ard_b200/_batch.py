"""Batched runs of many small NMF models over the C ABI's bnmtf_batch_run.

The reference fits the models of a cross-validation sweep one by one
(code/cross_validation/matrix_cross_validation.py:83-101) or in a multiprocessing.Pool
(parallel_matrix_cross_validation.py:53-66).  Here every model is constructed and initialised
as usual, but its run() only uploads the state (`deferred`); `run_deferred` then drives the
device loops of all of them together, one kernel launch per update for the whole batch, and
hands each model its own results.  A model run in a batch follows bit for bit the chain it
follows alone.
"""
import contextlib
import ctypes as C

import numpy as np

from . import _lib
from ._lib import NSTATS, check, dptr


def layout_floor(masks):
    ''' Packing maxima shared by the models of a batch: the longest observed row and column over all
        training masks (bnmtf_nmf_layout_stats order: rows, row tile segment, columns, column tile segment). '''
    rows = cols = 1
    for M in masks:
        M = np.asarray(M) != 0
        rows = max(rows, int(M.sum(axis=1).max()))
        cols = max(cols, int(M.sum(axis=0).max()))
    return np.array([rows, 1, cols, 1], dtype=np.int64)


def reserve_device_memory(n_models, I, J, n_observed, workers=1):
    ''' Grow the device memory pool once for n_models handles of an I x J matrix with n_observed entries:
        packed rows and columns (12 B per slot and side, padded) plus the dense blocks that `workers`
        concurrent constructors hold while packing. '''
    lib = _lib.require_device()
    per_model = 2 * 16 * int(n_observed) + 64 * (int(I) + int(J)) * 40
    transient = int(workers) * 20 * int(I) * int(J)
    check(lib.bnmtf_pool_reserve(_lib.current_device(), int(n_models) * per_model + transient))


@contextlib.contextmanager
def deferred(model, floor=None):
    ''' Inside this block model.run(iterations) -- also when called through train() -- uploads the
        state and returns; finish with run_deferred([...]). '''
    model._defer_run = True
    if floor is not None:
        model._layout_floor = floor
    try:
        yield model
    finally:
        model._defer_run = False


def supports_batching(model):
    return hasattr(model, '_run_begin') and hasattr(model, '_defer_run') and hasattr(model, 'METHOD')


last_device_seconds = []     # CUDA-event seconds of each bnmtf_batch_run of the last run_deferred call


def run_deferred(models):
    ''' Run the device loops of all models whose run() was deferred.  Models are grouped by
        (method, iterations); every group is one bnmtf_batch_run.  Returns the number of kernel
        launches. '''
    groups = {}
    for m in models:
        if m._deferred is None:
            raise ValueError("run_deferred: a model has no deferred run()")
        iterations, eng, seed, traces = m._deferred
        if traces:
            raise ValueError("Batched runs do not store the draws (all_U, all_V): call "
                             "summarise_on_device(burn_in, thinning) on Gibbs / ICM models before train().")
        groups.setdefault((m.METHOD, iterations), []).append(m)
    lib = _lib.require_device()
    total = 0
    del last_device_seconds[:]
    for (method, iterations), ms in groups.items():
        n = len(ms)
        handles = (C.c_void_p * n)(*[m._deferred[1].h for m in ms])
        seeds = (C.c_uint64 * n)(*[int(m._deferred[2]) for m in ms])
        stats = [np.zeros((iterations, NSTATS)) for _ in ms]
        stat_ptrs = (C.POINTER(C.c_double) * n)(*[dptr(s) for s in stats])
        seconds = np.zeros(iterations)
        launches = C.c_int64(0)
        check(lib.bnmtf_batch_run(n, handles, int(method), int(iterations), seeds, stat_ptrs, dptr(seconds),
                                  C.byref(launches)))
        total += int(launches.value)
        last_device_seconds.append(float(seconds[-1]) if iterations > 0 else 0.)
        for m, st in zip(ms, stats):
            _, eng, _, _ = m._deferred
            m._deferred = None
            m._run_end(iterations, eng, {"all_U": None, "all_V": None, "all_lambdak": None, "stats": st,
                                         "seconds": seconds.copy()})
    return total
