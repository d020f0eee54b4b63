"""Shared helpers: Philox seeds derived from numpy's global RandomState (so that
numpy.random.seed(s) keeps runs reproducible, SURVEY.md 8b 'RNG'), and the
vector calls into the C ABI."""
import ctypes as C

import numpy

from .. import _lib
from .._lib import check, dptr


def next_seed():
    """One 62-bit Philox key, consuming two draws of numpy's global legacy stream."""
    lo = int(numpy.random.randint(0, 2 ** 31 - 1))
    hi = int(numpy.random.randint(0, 2 ** 31 - 1))
    seed = (hi << 31) | lo
    # one process per GPU: every rank consumes its own numpy stream, but all of them must key the device
    # samplers identically (they replicate the lambda / tau draws and the other ranks' shards are conditioned on
    # them), so rank 0's seed is the seed of the job whether or not the caller seeded numpy alike on every rank
    from .. import _dist
    if _dist.info()[1] > 1:
        seed = _dist.broadcast_int(seed, src=0)
    return seed


def _vec(x):
    return numpy.ascontiguousarray(numpy.atleast_1d(numpy.asarray(x, dtype=numpy.float64)))


def tn_draw(mus, taus, seed=None):
    lib = _lib.require_device()
    mus, taus = _vec(mus), _vec(taus)
    out = numpy.empty_like(mus)
    check(lib.bnmtf_tn_draw(_lib.current_device(), C.c_uint64(next_seed() if seed is None else seed), mus.size,
                            dptr(mus), dptr(taus), dptr(out)))
    return out


def tn_moments(mus, taus):
    lib = _lib.require_device()
    mus, taus = _vec(mus), _vec(taus)
    e, v = numpy.empty_like(mus), numpy.empty_like(mus)
    check(lib.bnmtf_tn_moments(_lib.current_device(), mus.size, dptr(mus), dptr(taus), dptr(e), dptr(v)))
    return e, v


def gamma_draws(alphas, betas, seed=None):
    lib = _lib.require_device()
    alphas, betas = _vec(alphas), _vec(betas)
    out = numpy.empty_like(alphas)
    check(lib.bnmtf_gamma_draw(_lib.current_device(), C.c_uint64(next_seed() if seed is None else seed), alphas.size,
                               dptr(alphas), dptr(betas), dptr(out)))
    return out


def exp_draws(lambdas, seed=None):
    lib = _lib.require_device()
    lambdas = _vec(lambdas)
    out = numpy.empty_like(lambdas)
    check(lib.bnmtf_exp_draw(_lib.current_device(), C.c_uint64(next_seed() if seed is None else seed), lambdas.size,
                             dptr(lambdas), dptr(out)))
    return out
