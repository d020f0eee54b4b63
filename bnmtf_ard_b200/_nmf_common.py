"""Shared host-side logic of the four NMF classes: constructor checks with the
reference's exact messages, lazy engine creation, prediction metrics.  The
state (U, V, tau, ...) lives in plain numpy attributes on the object and is
host-authoritative between calls, exactly like the reference objects; every
method that computes uploads what it needs and downloads its results."""
import math

import numpy

from . import _lib
from ._engine import NMFEngine

ALL_METRICS = ['MSE', 'R^2', 'Rp']
ALL_QUALITY = ['loglikelihood', 'BIC', 'AIC', 'MSE', 'ELBO']


class NMFCommon(object):
    verbose = True      # print one line per iteration like the reference (after the device loop)

    def _setup_data(self, R, M, K):
        ''' Constructor checks shared by all models (bnmf_gibbs.py:61-75). '''
        self.R = numpy.array(R, dtype=float)
        self.M = numpy.array(M, dtype=float)
        self.K = K

        assert len(self.R.shape) == 2, "Input matrix R is not a two-dimensional array, " \
            "but instead %s-dimensional." % len(self.R.shape)
        assert self.R.shape == self.M.shape, "Input matrix R is not of the same size as " \
            "the indicator matrix M: %s and %s respectively." % (self.R.shape, self.M.shape)

        (self.I, self.J) = self.R.shape
        self.size_Omega = self.M.sum()
        self.check_empty_rows_columns()
        self._eng = None

    def _setup_hyper(self, ARD, hyperparameters):
        ''' Hyper-parameter handling of the Bayesian models (bnmf_gibbs.py:76-89). '''
        self.ARD = ARD
        self.alphatau, self.betatau = float(hyperparameters['alphatau']), float(hyperparameters['betatau'])
        if self.ARD:
            self.alpha0, self.beta0 = float(hyperparameters['alpha0']), float(hyperparameters['beta0'])
        else:
            self.lambdaU, self.lambdaV = numpy.array(hyperparameters['lambdaU']), numpy.array(hyperparameters['lambdaV'])
            # Make lambdaU/V into a numpy array if they are a float
            if self.lambdaU.shape == ():
                self.lambdaU = self.lambdaU * numpy.ones((self.I, self.K))
            if self.lambdaV.shape == ():
                self.lambdaV = self.lambdaV * numpy.ones((self.J, self.K))

            assert self.lambdaU.shape == (self.I, self.K), "Prior matrix lambdaU has the wrong shape: %s instead of (%s, %s)." % (self.lambdaU.shape, self.I, self.K)
            assert self.lambdaV.shape == (self.J, self.K), "Prior matrix lambdaV has the wrong shape: %s instead of (%s, %s)." % (self.lambdaV.shape, self.J, self.K)

    def check_empty_rows_columns(self):
        ''' Raise an exception if an entire row or column is empty (bnmf_gibbs.py:92-101). '''
        sums_columns = self.M.sum(axis=0)
        sums_rows = self.M.sum(axis=1)
        empty_rows = numpy.flatnonzero(sums_rows == 0)
        assert empty_rows.size == 0, "Fully unobserved row in R, row %s." % (empty_rows[0] if empty_rows.size else -1)
        empty_columns = numpy.flatnonzero(sums_columns == 0)
        assert empty_columns.size == 0, "Fully unobserved column in R, column %s." % (empty_columns[0] if empty_columns.size else -1)

    # -- engine ----------------------------------------------------------------
    _layout_floor = None    # batched runs: lower bounds of the packing maxima, so that all models share one layout
    _defer_run = False      # batched runs: run() uploads the state and stops; _batch.run_deferred() finishes it
    _deferred = None

    def _engine(self):
        ''' The device handle, created on first use (packs R, M on the GPU). '''
        if self._eng is None:
            self._eng = NMFEngine(self.R, self.M, self.K, layout_floor=self._layout_floor)
        return self._eng

    def run(self, iterations):
        ''' run(iterations) of the reference classes: upload the host state, iterate on the device,
            download.  The three steps are separate so that many small models can share the device
            loop (bnmtf_ard_b200._batch). '''
        eng, seed, traces = self._run_begin(iterations)
        if self._defer_run:
            self._deferred = (int(iterations), eng, seed, traces)
            return
        res = eng.run(self.METHOD, iterations, seed=seed, traces=traces, want_lambdak=self._WANT_LAMBDAK)
        self._run_end(iterations, eng, res)

    def _upload_hyper(self, eng):
        if getattr(self, 'ARD', True):
            eng.set_hyper(True, getattr(self, 'alphatau', 1.), getattr(self, 'betatau', 1.),
                          getattr(self, 'alpha0', 1.), getattr(self, 'beta0', 1.))
        else:
            eng.set_hyper(False, self.alphatau, self.betatau, 1., 1., self.lambdaU, self.lambdaV)

    def _fill_performances(self, eng, stats):
        perf = eng.performances_all(stats)      # all iterations at once (same expressions as eng.performances)
        for metric in ALL_METRICS:
            self.all_performances[metric].extend(perf[metric])

    # -- metrics (bnmf_gibbs.py:252-270; identical in all eight models) --------
    def compute_MSE(self, M, R, R_pred):
        ''' Return the MSE of predictions in R_pred, expected values in R, for the entries in M. '''
        return (M * (R - R_pred)**2).sum() / float(M.sum())

    def compute_R2(self, M, R, R_pred):
        ''' Return the R^2 of predictions in R_pred, expected values in R, for the entries in M. '''
        mean = (M * R).sum() / float(M.sum())
        SS_total = float((M * (R - mean)**2).sum())
        SS_res = float((M * (R - R_pred)**2).sum())
        return 1. - SS_res / SS_total if SS_total != 0. else numpy.inf

    def compute_Rp(self, M, R, R_pred):
        ''' Return the Rp of predictions in R_pred, expected values in R, for the entries in M. '''
        mean_real = (M * R).sum() / float(M.sum())
        mean_pred = (M * R_pred).sum() / float(M.sum())
        covariance = (M * (R - mean_real) * (R_pred - mean_pred)).sum()
        variance_real = (M * (R - mean_real)**2).sum()
        variance_pred = (M * (R_pred - mean_pred)**2).sum()
        return covariance / float(math.sqrt(variance_real) * math.sqrt(variance_pred))

    # -- on-device posterior summaries (opt-in, not in the reference API) ------
    _summary = None        # (burn_in, thinning) handed to summarise_on_device
    _keep_traces = True
    _summary_ready = False

    def summarise_on_device(self, burn_in, thinning, keep_traces=False):
        ''' Before run(): accumulate the posterior means over range(burn_in, iterations, thinning)
            on the GPU instead of storing every draw (all_U, all_V / all_F, all_S, all_G become None
            unless keep_traces).  approx_expectation / predict / quality called with the same
            (burn_in, thinning) then read the device summary.  Pass thinning=None to switch off. '''
        if thinning is None:
            self._summary, self._keep_traces, self._summary_ready = None, True, False
            return
        assert burn_in >= 0 and thinning >= 1, "Need burn_in >= 0 and thinning >= 1."
        self._summary, self._keep_traces, self._summary_ready = (int(burn_in), int(thinning)), bool(keep_traces), False

    def _configure_summary(self, eng):
        if self._summary is None:
            eng.set_summary(0, 0)
            return True
        eng.set_summary(*self._summary)
        return self._keep_traces

    def _summary_matches(self, burn_in, thinning):
        return self._summary_ready and self._summary == (int(burn_in), int(thinning))

    def _need_traces(self, traces):
        if traces is None:
            raise ValueError("The draws were not stored (summarise_on_device): only the (burn_in, thinning) = %s "
                             "summary is available." % (self._summary,))

    def _device_predict(self, M_pred, source, eng=None):
        ''' MSE, R^2, Rp over the entries of M_pred, evaluated on the GPU from the device factors. '''
        eng = self._engine() if eng is None else eng
        rows, cols = numpy.nonzero(numpy.asarray(M_pred))
        res = eng.predict_entries(source, rows, cols, self.R[rows, cols])
        self._last_predict = res
        return {'MSE': res['MSE'], 'R^2': res['R^2'], 'Rp': res['Rp']}

    def _shapes_fit_handle(self, **factors):
        ''' False when a factor was assigned with another number of columns than the constructor's K (L): the
            reference's own tests do (tests/code/test_bnmf_vb.py:368-393, test_nmf_np.py:222-243) and numpy.dot does
            not care, but the device handle is built for K (L) columns.  predict() then evaluates the three metrics
            the way the sampled classes' predict() does -- on the host, off the iteration loop. '''
        return all(numpy.shape(v) == want for v, want in factors.values())

    def _predict_from(self, M_pred, X, Y):
        R_pred = numpy.dot(X, Y.T)
        MSE = self.compute_MSE(M_pred, self.R, R_pred)
        R2 = self.compute_R2(M_pred, self.R, R_pred)
        Rp = self.compute_Rp(M_pred, self.R, R_pred)
        return {'MSE': MSE, 'R^2': R2, 'Rp': Rp}


NMTF_MAX_L, NMTF_MAX_KL = 48, 3631      # csrc/engine.cu: NMTF_MAX_L, NMTF_MAX_KL


def check_nmtf_limits(K, L):
    ''' Size limits of the tri-factorisation engine (the reference has none): raised in the constructor, with the
        same text the C ABI would give at handle creation, never in the middle of run(). '''
    if int(L) > NMTF_MAX_L:
        raise ValueError("tri-factorisation: L = %d is not supported by the B200 engine (L <= %d; K is only bound by "
                         "K*L <= %d)" % (L, NMTF_MAX_L, NMTF_MAX_KL))
    if int(K) * int(L) > NMTF_MAX_KL:
        raise ValueError("tri-factorisation: K*L = %d is not supported by the B200 engine (K*L <= %d)" % (int(K) * int(L), NMTF_MAX_KL))


def method_id(name):
    return {'gibbs': _lib.GIBBS, 'icm': _lib.ICM, 'vb': _lib.VB, 'np': _lib.NP}[name]
