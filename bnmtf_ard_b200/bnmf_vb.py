"""
Variational Bayesian inference for non-negative matrix factorisation, with ARD
-- B200 engine.  Drop-in for the reference class code/models/bnmf_vb.py: same
constructor (R, M, K, ARD, hyperparameters), initialise(init_UV), run(iterations),
train, predict(M_pred), quality(metric), the variational parameters as public
attributes (mu_U, tau_U, exp_U, var_U, ..V, alpha_s, beta_s, exp_tau, exp_logtau,
alphak_s, betak_s, exp_lambdak, exp_loglambdak) and the update_* methods.
"""
import math

import numpy

from . import _lib
from ._lib import F_MU, F_TAU, F_VALUE, F_VAR, SIDE_U, SIDE_V
from ._nmf_common import ALL_METRICS, ALL_QUALITY, NMFCommon
from .distributions._device import next_seed
from .distributions.gamma import gamma_expectation, gamma_expectation_log
from .distributions.truncated_normal_vector import TN_vector_expectation, TN_vector_variance

OPTIONS_INIT_UV = ['random', 'exp']


class bnmf_vb(NMFCommon):
    METHOD = _lib.VB

    def __init__(self, R, M, K, ARD, hyperparameters):
        ''' Set up the class and do some checks on the values passed. '''
        self._setup_data(R, M, K)
        self._setup_hyper(ARD, hyperparameters)

    def train(self, init_UV, iterations):
        ''' Initialise and run the algorithm. '''
        self.initialise(init_UV=init_UV)
        self.run(iterations)

    def initialise(self, init_UV='exp'):
        ''' Initialise U, V, tau, and lambda (if ARD).  (bnmf_vb.py:105-142) '''
        assert init_UV in OPTIONS_INIT_UV, "Unknown initialisation option: %s. Should be in %s." % (init_UV, OPTIONS_INIT_UV)

        # Initialise lambdak, and compute expectation
        if self.ARD:
            self.alphak_s, self.betak_s = numpy.zeros(self.K), numpy.zeros(self.K)
            self.exp_lambdak, self.exp_loglambdak = numpy.zeros(self.K), numpy.zeros(self.K)
            for k in range(self.K):
                self.alphak_s[k] = self.alpha0
                self.betak_s[k] = self.beta0
                self.update_exp_lambdak(k)

        # mu = Exp draw or 1/lambda, tau = 1, moments of TN(mu, 1): on the device
        eng = self._engine()
        self._upload_hyper(eng)
        eng.set_scalars(0., self.exp_lambdak if self.ARD else None)
        eng.init_factors(self.METHOD, init_UV == 'random', next_seed() if init_UV == 'random' else 0)
        self._download_factors(eng)

        # Initialise tau and compute expectation
        self.exp_tau, self.exp_logtau = 0., 0.
        self.alpha_s, self.beta_s = 0., 0.
        self.update_tau()
        self.update_exp_tau()

    def _download_factors(self, eng):
        self.mu_U, self.tau_U = eng.get_factor(SIDE_U, F_MU), eng.get_factor(SIDE_U, F_TAU)
        self.exp_U, self.var_U = eng.get_factor(SIDE_U, F_VALUE), eng.get_factor(SIDE_U, F_VAR)
        self.mu_V, self.tau_V = eng.get_factor(SIDE_V, F_MU), eng.get_factor(SIDE_V, F_TAU)
        self.exp_V, self.var_V = eng.get_factor(SIDE_V, F_VALUE), eng.get_factor(SIDE_V, F_VAR)

    def _upload_state(self):
        eng = self._engine()
        self._upload_hyper(eng)
        # The state lives in plain attributes that callers -- the reference's tests above all
        # (tests/code/test_bnmf_vb.py:231-290) -- assign one by one before calling a single update; what has not been
        # assigned yet does not enter that update and goes up as a neutral value (mu 0, tau 1, moments 0, scalars 1 / 0).
        for side, X, n in ((SIDE_U, 'U', self.I), (SIDE_V, 'V', self.J)):
            eng.set_factor(side, F_MU, self._state('mu_' + X, (n, self.K), 0.))
            eng.set_factor(side, F_TAU, self._state('tau_' + X, (n, self.K), 1.))
            eng.set_factor(side, F_VALUE, self._state('exp_' + X, (n, self.K), 0.))
            eng.set_factor(side, F_VAR, self._state('var_' + X, (n, self.K), 0.))
        sc = [self._state('exp_tau', (), 1.), self._state('exp_logtau', (), 0.), self._state('alpha_s', (), 1.), self._state('beta_s', (), 1.)]
        if self.ARD:
            eng.set_vb_scalars(*sc, self._state('alphak_s', (self.K,), 1.), self._state('betak_s', (self.K,), 1.),
                               self._state('exp_lambdak', (self.K,), 1.), self._state('exp_loglambdak', (self.K,), 0.))
        else:
            eng.set_vb_scalars(*sc)
        return eng

    def _state(self, name, shape, default):
        v = getattr(self, name, None)
        if v is None or callable(v):        # (alpha_s / beta_s are methods of the sampled parent class until VB assigns them)
            return default if shape == () else numpy.full(shape, default)
        return v

    _WANT_LAMBDAK = False

    def _run_begin(self, iterations):
        ''' Run the VB updates (bnmf_vb.py:145-188): reset the traces, upload the state. '''
        self.all_exp_tau = []  # to check for convergence
        self.all_times = []  # to plot performance against time
        self.all_performances = {}  # for plotting convergence of metrics
        for metric in ALL_METRICS:
            self.all_performances[metric] = []
        return self._upload_state(), 0, False

    def _run_end(self, iterations, eng, res):
        self._download_factors(eng)
        sc, ls = eng.get_vb_scalars()
        if iterations > 0:
            self.exp_tau, self.exp_logtau, self.alpha_s, self.beta_s = [float(v) for v in sc]
            if self.ARD:
                self.alphak_s, self.betak_s = ls[0].copy(), ls[1].copy()
                self.exp_lambdak, self.exp_loglambdak = ls[2].copy(), ls[3].copy()
        self.all_exp_tau = [float(v) for v in res['stats'][:, _lib.ST_TAU]]
        self.all_elbo = [float(v) for v in res['stats'][:, _lib.ST_ELBO]]
        self._fill_performances(eng, res['stats'])
        self.all_times = list(res['seconds'])
        if self.verbose:
            for it in range(iterations):
                print("Iteration %s. ELBO: %s. MSE: %s. R^2: %s. Rp: %s." % (
                    it + 1, self.all_elbo[it], self.all_performances['MSE'][it], self.all_performances['R^2'][it],
                    self.all_performances['Rp'][it]))

    def elbo(self):
        ''' Compute the ELBO.  (bnmf_vb.py:191-232) '''
        return float(self._upload_state().stats(self.METHOD)[_lib.ST_ELBO])

    ''' Update the parameters for the distributions. '''
    def update_tau(self):
        ''' Parameter updates tau. '''
        self.alpha_s = self.alphatau + self.size_Omega / 2.0
        self.beta_s = self.betatau + 0.5 * self.exp_square_diff()

    def exp_square_diff(self):
        ''' Compute: sum_Omega E_q(U,V) [ ( Rij - Ui Vj )^2 ].  (bnmf_vb.py:241-244) '''
        return float(self._upload_state().stats(self.METHOD)[_lib.ST_ESD])

    def update_lambdak(self, k):
        ''' Parameter updates lambdak. '''
        self.alphak_s[k] = self.alpha0 + self.I + self.J
        self.betak_s[k] = self.beta0 + self.exp_U[:, k].sum() + self.exp_V[:, k].sum()

    def update_U(self, k):
        ''' Parameter updates U.  (bnmf_vb.py:251-255) '''
        self.tau_U[:, k], self.mu_U[:, k] = self._upload_state().column_params(self.METHOD, SIDE_U, k)

    def update_V(self, k):
        ''' Parameter updates V.  (bnmf_vb.py:257-261) '''
        self.tau_V[:, k], self.mu_V[:, k] = self._upload_state().column_params(self.METHOD, SIDE_V, k)

    ''' Update the expectations and variances. '''
    def update_exp_tau(self):
        ''' Update expectation tau. '''
        self.exp_tau = gamma_expectation(self.alpha_s, self.beta_s)
        self.exp_logtau = gamma_expectation_log(self.alpha_s, self.beta_s)

    def update_exp_lambdak(self, k):
        ''' Update expectation lambdak. '''
        self.exp_lambdak[k] = gamma_expectation(self.alphak_s[k], self.betak_s[k])
        self.exp_loglambdak[k] = gamma_expectation_log(self.alphak_s[k], self.betak_s[k])

    def update_exp_U(self, k):
        ''' Update expectation U. '''
        self.exp_U[:, k] = TN_vector_expectation(self.mu_U[:, k], self.tau_U[:, k])
        self.var_U[:, k] = TN_vector_variance(self.mu_U[:, k], self.tau_U[:, k])

    def update_exp_V(self, k):
        ''' Update expectation V. '''
        self.exp_V[:, k] = TN_vector_expectation(self.mu_V[:, k], self.tau_V[:, k])
        self.var_V[:, k] = TN_vector_variance(self.mu_V[:, k], self.tau_V[:, k])

    def predict(self, M_pred):
        ''' Predict missing values in R. '''
        if not self._shapes_fit_handle(U=(self.exp_U, (self.I, self.K)), V=(self.exp_V, (self.J, self.K))):
            return self._predict_from(M_pred, numpy.asarray(self.exp_U, dtype=float), numpy.asarray(self.exp_V, dtype=float))
        return self._device_predict(M_pred, _lib.PRED_CURRENT, self._upload_state())

    def quality(self, metric):
        ''' Return the model quality, either as log likelihood, BIC, AIC, MSE, or ELBO. '''
        assert metric in ALL_QUALITY, 'Unrecognised metric for model quality: %s.' % metric

        log_likelihood = self.log_likelihood()
        if metric == 'loglikelihood':
            return log_likelihood
        elif metric == 'BIC':
            # -2*loglikelihood + (no. free parameters * log(no data points))
            return - 2 * log_likelihood + self.number_parameters() * math.log(self.size_Omega)
        elif metric == 'AIC':
            # -2*loglikelihood + 2*no. free parameters
            return - 2 * log_likelihood + 2 * self.number_parameters()
        elif metric == 'MSE':
            R_pred = numpy.dot(self.exp_U, self.exp_V.T)
            return self.compute_MSE(self.M, self.R, R_pred)
        elif metric == 'ELBO':
            return self.elbo()

    def log_likelihood(self):
        ''' Return the likelihood of the data given the trained model's parameters. '''
        return self.size_Omega / 2. * (self.exp_logtau - math.log(2 * math.pi)) \
            - self.exp_tau / 2. * (self.M * (self.R - numpy.dot(self.exp_U, self.exp_V.T))**2).sum()

    def number_parameters(self):
        ''' Return the number of free variables in the model. '''
        return (self.I * self.K + self.J * self.K + 1) + (self.K if self.ARD else 0)
