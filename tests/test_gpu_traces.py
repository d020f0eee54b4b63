"""GPU: the way of the draw traces (all_U, all_V) to the caller's arrays.  Long runs stage them on the device in
chunks, copy every chunk into a pinned host slot and let a worker thread move it into the caller's pageable arrays
(TraceDrain in engine.cu); short runs copy one chunk at the end.  All ways must deliver the same bits."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HP = {'alphatau': 1., 'betatau': 1., 'alpha0': 1., 'beta0': 1.}


def problem(I=300, J=200, K=5, seed=0):
    rng = np.random.default_rng(seed)
    R = rng.exponential(1, (I, K)) @ rng.exponential(1, (J, K)).T + rng.normal(0, 0.5, (I, J))
    M = (rng.random((I, J)) < 0.6).astype(float)
    M[np.arange(I), rng.integers(0, J, I)] = 1
    M[rng.integers(0, I, J), np.arange(J)] = 1
    return R, M


@pytest.mark.parametrize("iterations", [37, 64])
def test_chunked_traces_equal_single_chunk(monkeypatch, iterations):
    from bnmtf_ard_b200 import bnmf_gibbs
    R, M = problem()
    out = []
    # per iteration (300 + 200) * 6 * 8 = 24 KB of draws: 200 KB chunks hold 8 iterations
    for env in ({}, {"BNMTF_TRACE_CHUNK_BYTES": "200000"}, {"BNMTF_TRACE_CHUNK_BYTES": "200000", "BNMTF_NO_PINNED_TRACES": "1"},
                {"BNMTF_TRACE_CHUNK_BYTES": "30000"}):
        for k in ("BNMTF_TRACE_CHUNK_BYTES", "BNMTF_NO_PINNED_TRACES"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        np.random.seed(3)
        m = bnmf_gibbs(R, M, 6, True, HP); m.verbose = False
        m.train(init_UV='random', iterations=iterations)
        out.append((m.all_U.copy(), m.all_V.copy(), np.array(m.all_tau), np.array(m.all_lambdak)))
        # a second run on the same handle reuses the pinned slots
        m.run(5)
        assert m.all_U.shape == (5, 300, 6) and np.isfinite(m.all_U).all()
    for other in out[1:]:
        for a, b in zip(out[0], other):
            assert np.array_equal(a, b)
    assert (out[0][0][-1] != out[0][0][0]).any()


def test_chunked_traces_nmtf(monkeypatch):
    from bnmtf_ard_b200 import bnmtf_gibbs
    R, M = problem(120, 90, 3, 2)
    hp = dict(HP, lambdaS=0.1)
    out = []
    for env in ({}, {"BNMTF_TRACE_CHUNK_BYTES": "60000"}, {"BNMTF_TRACE_CHUNK_BYTES": "60000", "BNMTF_NO_PINNED_TRACES": "1"}):
        for k in ("BNMTF_TRACE_CHUNK_BYTES", "BNMTF_NO_PINNED_TRACES"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        np.random.seed(5)
        m = bnmtf_gibbs(R, M, 4, 3, True, hp); m.verbose = False
        m.train(init_FG='random', init_S='random', iterations=30)
        out.append((np.array(m.all_F), np.array(m.all_S), np.array(m.all_G), np.array(m.all_tau)))
    for other in out[1:]:
        for a, b in zip(out[0], other):
            assert np.array_equal(a, b)
