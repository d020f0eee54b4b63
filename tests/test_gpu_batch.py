"""GPU: batched runs of many small models (bnmtf_batch_run).

The batched kernels are the bodies of the single-model kernels with blockIdx.y selecting the model,
so the bar is bit-exactness: a model run in a batch must end in exactly the state -- and report
exactly the per-iteration statistics -- of the same model run alone."""
import contextlib
import io

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HP = {'alphatau': 1., 'betatau': 1., 'alpha0': 1., 'beta0': 1.}


def problem(I=70, J=50, K=4, seed=0, rho=0.8):
    rng = np.random.default_rng(seed)
    R = rng.exponential(1, (I, K)) @ rng.exponential(1, (J, K)).T + rng.normal(0, 0.5, (I, J))
    M = (rng.random((I, J)) < rho).astype(float)
    M[np.arange(I), rng.integers(0, J, I)] = 1
    M[rng.integers(0, I, J), np.arange(J)] = 1
    return R, M


def make_models(name, R, masks, Ks, floor, ARD=True):
    import bnmtf_ard_b200 as pkg
    cls = getattr(pkg, name)
    models = []
    for M, K in zip(masks, Ks):
        if name == 'nmf_np':
            m = cls(np.abs(R), M, K)
        elif ARD:
            m = cls(R, M, K, True, HP)
        else:
            m = cls(R, M, K, False, {'alphatau': 1., 'betatau': 1., 'lambdaU': 0.7, 'lambdaV': 1.3})
        m.verbose = False
        m._layout_floor = floor
        models.append(m)
    return models


def state_of(m):
    names = ('U', 'V', 'tau', 'lambdak', 'exp_U', 'exp_V', 'var_U', 'var_V', 'mu_U', 'tau_U', 'mu_V', 'tau_V', 'exp_tau',
             'exp_logtau', 'alpha_s', 'beta_s', 'alphak_s', 'betak_s', 'exp_lambdak', 'exp_loglambdak', 'all_elbo', 'all_i_div',
             'all_exp_tau', 'all_tau')
    out = {n: np.array(getattr(m, n)) for n in names if hasattr(m, n) and not callable(getattr(m, n)) and getattr(m, n) is not None}
    for k, v in m.all_performances.items():
        out['perf_' + k] = np.array(v)
    return out


@pytest.mark.parametrize("name,init,ARD", [('bnmf_vb', 'random', True), ('bnmf_vb', 'exp', False), ('nmf_icm', 'random', True),
                                           ('nmf_np', 'random', True), ('bnmf_gibbs', 'random', True), ('bnmf_gibbs', 'random', False)])
@pytest.mark.parametrize("iterations", [3, 25])     # 3: eager loop, 25: graph replays
def test_batch_equals_single_runs_bitwise(name, init, ARD, iterations):
    from bnmtf_ard_b200 import _batch
    R, M = problem()
    rng = np.random.default_rng(3)
    masks, Ks = [], [1, 2, 3, 5, 8, 4, 2]
    for _ in Ks:                                       # different training masks, like the folds of a CV
        T = M * (rng.random(M.shape) < 0.85)
        T[np.arange(M.shape[0]), np.argmax(M, axis=1)] = 1
        T[np.argmax(M, axis=0), np.arange(M.shape[1])] = 1
        masks.append(T)
    floor = _batch.layout_floor(masks)
    gibbs = name in ('bnmf_gibbs', 'nmf_icm')     # the classes that store their draws unless summarised on the device
    results = []
    for batched in (False, True):
        np.random.seed(11)
        models = make_models(name, R, masks, Ks, floor, ARD)
        for m in models:
            if gibbs:
                m.summarise_on_device(iterations // 3, 2)
            if batched:
                with _batch.deferred(m):
                    m.train(iterations=iterations, init_UV=init)
            else:
                m.train(iterations=iterations, init_UV=init)
        if batched:
            assert all(m._deferred is not None for m in models)
            launches = _batch.run_deferred(models)
            assert launches > 0 and all(m._deferred is None for m in models)
        res = [state_of(m) for m in models]
        if gibbs:
            for m, r in zip(models, res):
                eU, eV, etau, elam = m.approx_expectation(iterations // 3, 2)
                r.update(sum_U=eU, sum_V=eV, sum_tau=np.array(etau))
                if ARD:
                    r['sum_lam'] = elam
        results.append(res)
    for a, b in zip(*results):
        assert set(a) >= {'perf_MSE'} and set(a) - {'all_lambdak'} == set(b) - {'all_lambdak'}
        for key in b:
            assert np.array_equal(a[key], b[key], equal_nan=True), key


def test_batch_models_of_different_shapes():
    """The models of a batch only have to share the row-team layout, not their shapes."""
    from bnmtf_ard_b200 import _batch, bnmf_vb
    probs = [problem(40, 30, 3, 1), problem(55, 21, 2, 2), problem(33, 47, 4, 3)]
    floor = _batch.layout_floor([M for _, M in probs])
    out = []
    for batched in (False, True):
        np.random.seed(5)
        models = []
        for (R, M), K in zip(probs, (3, 2, 6)):
            m = bnmf_vb(R, M, K, True, HP); m.verbose = False; m._layout_floor = floor
            if batched:
                with _batch.deferred(m):
                    m.train(init_UV='random', iterations=12)
            else:
                m.train(init_UV='random', iterations=12)
            models.append(m)
        if batched:
            _batch.run_deferred(models)
        out.append([(m.exp_U.copy(), m.exp_V.copy(), np.array(m.all_elbo)) for m in models])
    for a, b in zip(*out):
        for x, y in zip(a, b):
            assert np.array_equal(x, y)


def test_batch_rejects_bad_input():
    from bnmtf_ard_b200 import _batch, bnmf_gibbs, bnmf_vb
    R, M = problem(30, 20, 2, 4)
    m = bnmf_gibbs(R, M, 2, True, HP); m.verbose = False
    with _batch.deferred(m):
        m.train(init_UV='random', iterations=5)
    with pytest.raises(ValueError, match="summarise_on_device"):     # no traces in a batch
        _batch.run_deferred([m])
    big = problem(30, 300, 2, 5, rho=1.0)
    a = bnmf_vb(R, M, 2, True, HP); b = bnmf_vb(big[0], big[1], 2, True, HP)
    for x in (a, b):                                                  # packed without a common floor
        x.verbose = False
        with _batch.deferred(x):
            x.train(init_UV='exp', iterations=5)
    with pytest.raises(RuntimeError, match="different row-team layout"):
        _batch.run_deferred([a, b])


class QuietVB(object):
    def __new__(cls, R, M, **parameters):
        from bnmtf_ard_b200 import bnmf_vb
        m = bnmf_vb(R, M, **parameters)
        m.verbose = False
        return m


class QuietGibbs(object):
    def __new__(cls, R, M, **parameters):
        from bnmtf_ard_b200 import bnmf_gibbs
        m = bnmf_gibbs(R, M, **parameters)
        m.verbose = False
        return m


@pytest.mark.parametrize("method,train_config,predict_config", [
    (QuietVB, {'iterations': 30, 'init_UV': 'random'}, {}),
    (QuietGibbs, {'iterations': 30, 'init_UV': 'random'}, {'burn_in': 10, 'thinning': 2}),
])
def test_batched_cv_equals_serial_cv(tmp_path, method, train_config, predict_config):
    """Rows and columns short enough that every fold packs into the same layout with or without the
    batch's common floor: the batched driver must then write the serial driver's log byte for byte
    (same folds, same seeds, same chains)."""
    from bnmtf_ard_b200.cross_validation.batched_matrix_cross_validation import BatchedMatrixCrossValidation
    from bnmtf_ard_b200.cross_validation.matrix_cross_validation import MatrixCrossValidation
    R, M = problem(44, 28, 3, 6)
    search = [{'K': k, 'ARD': True, 'hyperparameters': HP} for k in (1, 2, 3, 5)]
    out = []
    for cls in (MatrixCrossValidation, BatchedMatrixCrossValidation):
        np.random.seed(4)
        import random
        random.seed(9)
        f = str(tmp_path / (cls.__name__ + ".txt"))
        with contextlib.redirect_stdout(io.StringIO()):
            cv = cls(method=method, R=R, M=M, K=4, parameter_search=search, train_config=train_config,
                     predict_config=predict_config, file_performance=f)
            cv.run()
            best, perf = cv.find_best_parameters('MSE', True)
        out.append((cv.performances['MSE'], best['K'], open(f).read()))
    assert out[0] == out[1]
    assert "Best performances:" in out[1][2] and "exception" not in out[1][2]


def test_nested_cv_batched(tmp_path):
    from bnmtf_ard_b200.cross_validation.nested_matrix_cross_validation import MatrixNestedCrossValidation
    R, M = problem(44, 28, 2, 7)
    logs = []
    for mode in (False, 'batched'):
        np.random.seed(2)
        import random
        random.seed(3)
        d = tmp_path / str(mode); d.mkdir()
        with contextlib.redirect_stdout(io.StringIO()):
            cv = MatrixNestedCrossValidation(method=QuietVB, R=R, M=M, K=3, P=2,
                                             parameter_search=[{'K': k, 'ARD': True, 'hyperparameters': HP} for k in (1, 2, 4)],
                                             train_config={'iterations': 20, 'init_UV': 'random'}, predict_config={},
                                             file_performance=str(d / "o.txt"),
                                             files_nested_performances=[str(d / ("n%d.txt" % i)) for i in range(3)])
            cv.run(parallel=mode)
        logs.append([open(str(d / "o.txt")).read()] + [open(str(d / ("n%d.txt" % i))).read() for i in range(3)])
    assert logs[0] == logs[1] and logs[0][0].startswith("Average performances:")
