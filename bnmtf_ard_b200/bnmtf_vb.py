"""
Variational Bayesian inference for non-negative matrix tri-factorisation
(R ~ F S G^T), with ARD -- B200 engine.  Drop-in for the reference class
code/models/bnmtf_vb.py: constructor (R, M, K, L, ARD, hyperparameters),
initialise(init_FG, init_S), run(iterations), train, predict(M_pred),
quality(metric), the variational parameters as public attributes (mu_/tau_/exp_/var_
of F, S, G; alpha_s, beta_s, exp_tau, exp_logtau; alphaFk_s, betaFk_s, exp_lambdaFk,
exp_loglambdaFk and the Gl block) and the update_* methods.
"""
import math

import numpy

from . import _lib
from ._engine import NMTFEngine
from ._lib import F_MU, F_TAU, F_VALUE, F_VAR, SIDE_U, SIDE_V
from ._nmf_common import ALL_METRICS, ALL_QUALITY
from .bnmtf_gibbs import OPTIONS_INIT_FG, OPTIONS_INIT_S, bnmtf_gibbs
from .distributions._device import next_seed
from .distributions.gamma import gamma_expectation, gamma_expectation_log
from .distributions.truncated_normal import TN_expectation, TN_variance
from .distributions.truncated_normal_vector import TN_vector_expectation, TN_vector_variance

ELEMENT_WISE_SPARSITY = True  # bnmtf_vb.py:60 (unused there as well)


class bnmtf_vb(bnmtf_gibbs):
    METHOD = _lib.VB
    # constructor, checks and hyper-parameter handling are those of bnmtf_gibbs (bnmtf_vb.py:62-96 is identical)

    def initialise(self, init_FG='random', init_S='random'):
        ''' Initialise F, S, G, tau, and lambdaFk, lambdaGl (if ARD).  (bnmtf_vb.py:117-190) '''
        assert init_FG in OPTIONS_INIT_FG, "Unknown initialisation option for F and G: %s. Should be in %s." % (init_FG, OPTIONS_INIT_FG)
        assert init_S in OPTIONS_INIT_S, "Unknown initialisation option for S: %s. Should be in %s." % (init_S, OPTIONS_INIT_S)
        # Initialise lambdaFk, lambdaGl, and compute expectations
        if self.ARD:
            self.alphaFk_s, self.betaFk_s = numpy.zeros(self.K), numpy.zeros(self.K)
            self.alphaGl_s, self.betaGl_s = numpy.zeros(self.L), numpy.zeros(self.L)
            self.exp_lambdaFk, self.exp_loglambdaFk = numpy.zeros(self.K), numpy.zeros(self.K)
            self.exp_lambdaGl, self.exp_loglambdaGl = numpy.zeros(self.L), numpy.zeros(self.L)
            for k in range(self.K):
                self.alphaFk_s[k] = self.alpha0
                self.betaFk_s[k] = self.beta0
                self.update_exp_lambdaFk(k)
            for l in range(self.L):
                self.alphaGl_s[l] = self.alpha0
                self.betaGl_s[l] = self.beta0
                self.update_exp_lambdaGl(l)

        # mu = Exp draw or 1/lambda, tau = 1, moments of TN(mu, 1) for F, G and S: on the device
        eng = self._engine()
        self._upload_hyper(eng)
        eng.set_scalars(0., self.exp_lambdaFk if self.ARD else None, self.exp_lambdaGl if self.ARD else None)
        eng.init(self.METHOD, init_FG == 'random', init_S == 'random',
                 next_seed() if init_FG == 'random' else 0, next_seed() if init_S == 'random' else 0)
        self._download_factors(eng)
        if init_FG == 'kmeans':
            # bnmtf_vb.py:142-159: mu = one-hot cluster indicators, tau = 1, then the moments (:178-183)
            self.mu_F, self.mu_G = self._kmeans_FG(eng)
            self.tau_F, self.tau_G = numpy.ones((self.I, self.K)), numpy.ones((self.J, self.L))
            for k in range(self.K):
                self.update_exp_F(k)
            for l in range(self.L):
                self.update_exp_G(l)

        # Initialise tau and compute expectation
        self.exp_tau, self.exp_logtau = 0., 0.
        self.alpha_s, self.beta_s = 0., 0.
        self.update_tau()
        self.update_exp_tau()

    def _download_factors(self, eng):
        for side, X in ((SIDE_U, 'F'), (SIDE_V, 'G')):
            setattr(self, 'mu_' + X, eng.get_factor(side, F_MU)); setattr(self, 'tau_' + X, eng.get_factor(side, F_TAU))
            setattr(self, 'exp_' + X, eng.get_factor(side, F_VALUE)); setattr(self, 'var_' + X, eng.get_factor(side, F_VAR))
        self.mu_S, self.tau_S = eng.get_S(F_MU), eng.get_S(F_TAU)
        self.exp_S, self.var_S = eng.get_S(F_VALUE), eng.get_S(F_VAR)

    def _upload_state(self):
        eng = self._engine()
        self._upload_hyper(eng)
        # attributes that the caller has not assigned yet (the reference's tests set only what one update reads,
        # tests/code/test_bnmtf_vb.py:349-477) go up as neutral values: mu 0, tau 1, moments 0, scalars 1 / 0
        st = self._state
        for side, X, n, k in ((SIDE_U, 'F', self.I, self.K), (SIDE_V, 'G', self.J, self.L)):
            eng.set_factor(side, F_MU, st('mu_' + X, (n, k), 0.)); eng.set_factor(side, F_TAU, st('tau_' + X, (n, k), 1.))
            eng.set_factor(side, F_VALUE, st('exp_' + X, (n, k), 0.)); eng.set_factor(side, F_VAR, st('var_' + X, (n, k), 0.))
        kl = (self.K, self.L)
        eng.set_S(F_MU, st('mu_S', kl, 0.)); eng.set_S(F_TAU, st('tau_S', kl, 1.))
        eng.set_S(F_VALUE, st('exp_S', kl, 0.)); eng.set_S(F_VAR, st('var_S', kl, 0.))
        sc = [st('exp_tau', (), 1.), st('exp_logtau', (), 0.), st('alpha_s', (), 1.), st('beta_s', (), 1.)]
        if self.ARD:
            eng.set_vb_scalars(*sc,
                               numpy.array([st('alphaFk_s', (self.K,), 1.), st('betaFk_s', (self.K,), 1.), st('exp_lambdaFk', (self.K,), 1.), st('exp_loglambdaFk', (self.K,), 0.)]),
                               numpy.array([st('alphaGl_s', (self.L,), 1.), st('betaGl_s', (self.L,), 1.), st('exp_lambdaGl', (self.L,), 1.), st('exp_loglambdaGl', (self.L,), 0.)]))
        else:
            eng.set_vb_scalars(*sc)
        return eng

    def _state(self, name, shape, default):
        v = getattr(self, name, None)
        if v is None or callable(v):        # (alpha_s / beta_s are methods of the sampled parent class until VB assigns them)
            return default if shape == () else numpy.full(shape, default)
        return v

    def run(self, iterations):
        ''' Run the VB updates.  (bnmtf_vb.py:193-245) '''
        self.all_exp_tau = []  # to check for convergence
        self.all_times = []  # to plot performance against time
        self.all_performances = {}  # for plotting convergence of metrics
        for metric in ALL_METRICS:
            self.all_performances[metric] = []

        eng = self._upload_state()
        res = eng.vb_run(iterations)
        self._download_factors(eng)
        sc, sf, sg = eng.get_vb_scalars()
        if iterations > 0:
            self.exp_tau, self.exp_logtau, self.alpha_s, self.beta_s = [float(v) for v in sc]
            if self.ARD:
                self.alphaFk_s, self.betaFk_s, self.exp_lambdaFk, self.exp_loglambdaFk = [sf[q].copy() for q in range(4)]
                self.alphaGl_s, self.betaGl_s, self.exp_lambdaGl, self.exp_loglambdaGl = [sg[q].copy() for q in range(4)]
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
        ''' Compute the ELBO.  (bnmtf_vb.py:248-299) '''
        return float(self._upload_state().vb_stats()[_lib.ST_ELBO])

    ''' Update the parameters for the distributions. '''
    def update_tau(self):
        ''' Parameter updates tau. '''
        self.alpha_s = self.alphatau + self.size_Omega / 2.0
        self.beta_s = self.betatau + 0.5 * self.exp_square_diff()

    def exp_square_diff(self):
        ''' Compute: sum_Omega E_q(F,S,G) [ ( Rij - Fi S Gj )^2 ].  (bnmtf_vb.py:319-324) '''
        return float(self._upload_state().vb_stats()[_lib.ST_ESD])

    def update_lambdaFk(self, k):
        ''' Parameter updates lambdaFk. '''
        self.alphaFk_s[k] = self.alpha0 + self.I
        self.betaFk_s[k] = self.beta0 + self.exp_F[:, k].sum()

    def update_lambdaGl(self, l):
        ''' Parameter updates lambdaGl. '''
        self.alphaGl_s[l] = self.alpha0 + self.J
        self.betaGl_s[l] = self.beta0 + self.exp_G[:, l].sum()

    def update_F(self, k):
        ''' Parameter updates F.  (bnmtf_vb.py:336-348) '''
        self.tau_F[:, k], self.mu_F[:, k] = self._upload_state().vb_update_params(0, k=k)

    def update_S(self, k, l):
        ''' Parameter updates S.  (bnmtf_vb.py:350-362) '''
        t, m = self._upload_state().vb_update_params(2, k=k, l=l)
        self.tau_S[k, l], self.mu_S[k, l] = float(t[0]), float(m[0])

    def update_G(self, l):
        ''' Parameter updates G.  (bnmtf_vb.py:364-375) '''
        self.tau_G[:, l], self.mu_G[:, l] = self._upload_state().vb_update_params(1, l=l)

    ''' Update the expectations and variances. '''
    def update_exp_tau(self):
        ''' Update expectation tau. '''
        self.exp_tau = gamma_expectation(self.alpha_s, self.beta_s)
        self.exp_logtau = gamma_expectation_log(self.alpha_s, self.beta_s)

    def update_exp_lambdaFk(self, k):
        ''' Update expectation lambdaFk. '''
        self.exp_lambdaFk[k] = gamma_expectation(self.alphaFk_s[k], self.betaFk_s[k])
        self.exp_loglambdaFk[k] = gamma_expectation_log(self.alphaFk_s[k], self.betaFk_s[k])

    def update_exp_lambdaGl(self, l):
        ''' Update expectation lambdaGl. '''
        self.exp_lambdaGl[l] = gamma_expectation(self.alphaGl_s[l], self.betaGl_s[l])
        self.exp_loglambdaGl[l] = gamma_expectation_log(self.alphaGl_s[l], self.betaGl_s[l])

    def update_exp_F(self, k):
        ''' Update expectation F. '''
        self.exp_F[:, k] = TN_vector_expectation(self.mu_F[:, k], self.tau_F[:, k])
        self.var_F[:, k] = TN_vector_variance(self.mu_F[:, k], self.tau_F[:, k])

    def update_exp_S(self, k, l):
        ''' Update expectation S. '''
        self.exp_S[k, l] = TN_expectation(self.mu_S[k, l], self.tau_S[k, l])
        self.var_S[k, l] = TN_variance(self.mu_S[k, l], self.tau_S[k, l])

    def update_exp_G(self, l):
        ''' Update expectation G. '''
        self.exp_G[:, l] = TN_vector_expectation(self.mu_G[:, l], self.tau_G[:, l])
        self.var_G[:, l] = TN_vector_variance(self.mu_G[:, l], self.tau_G[:, l])

    def predict(self, M_pred):
        ''' Predict missing values in R. '''
        if not self._shapes_fit_handle(F=(self.exp_F, (self.I, self.K)), S=(self.exp_S, (self.K, self.L)), G=(self.exp_G, (self.J, self.L))):
            FS = numpy.dot(numpy.asarray(self.exp_F, dtype=float), numpy.asarray(self.exp_S, dtype=float))
            return self._predict_from(M_pred, FS, numpy.asarray(self.exp_G, dtype=float))
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
            R_pred = self.triple_dot(self.exp_F, self.exp_S, self.exp_G.T)
            return self.compute_MSE(self.M, self.R, R_pred)
        elif metric == 'ELBO':
            return self.elbo()

    def log_likelihood(self):
        ''' Return the likelihood of the data given the trained model's parameters. '''
        return self.size_Omega / 2. * (self.exp_logtau - math.log(2 * math.pi)) \
            - self.exp_tau / 2. * (self.M * (self.R - self.triple_dot(self.exp_F, self.exp_S, self.exp_G.T))**2).sum()

    # Gibbs-only entry points of the parent that do not exist on the reference's bnmtf_vb
    tauF = muF = tauS = muS = tauG = muG = approx_expectation = predict_while_running = None
