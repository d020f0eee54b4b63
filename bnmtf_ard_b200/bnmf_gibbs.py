"""
Gibbs sampler for non-negative matrix factorisation, with ARD -- B200 engine.

Drop-in for the reference class code/models/bnmf_gibbs.py: same constructor
(R, M, K, ARD, hyperparameters), initialise(init_UV), run(iterations),
train(init_UV, iterations), predict(M_pred, burn_in, thinning), quality(...),
the same public attributes (U, V, tau, lambdak, all_U, all_V, all_tau,
all_lambdak, all_times, all_performances) and the per-parameter methods its
tests call.  The inference loop runs on the GPU (bnmtf_ard_b200/csrc); there is
no CPU fallback.

Usage of class:
    BNMF = bnmf_gibbs(R,M,K,ARD,hyperparameters)
    BNMF.initialise(init_UV)
    BNMF.run(iterations)
Or:
    BNMF = bnmf_gibbs(R,M,K,ARD,hyperparameters)
    BNMF.train(init_UV,iterations)
"""
import math

import numpy

from . import _lib
from ._lib import F_VALUE, SIDE_U, SIDE_V
from ._nmf_common import ALL_METRICS, ALL_QUALITY, NMFCommon
from .distributions._device import next_seed
from .distributions.gamma import gamma_draw, gamma_mode

OPTIONS_INIT_UV = ['random', 'exp']


class bnmf_gibbs(NMFCommon):
    METHOD = _lib.GIBBS

    def __init__(self, R, M, K, ARD, hyperparameters):
        ''' Set up the class and do some checks on the values passed. '''
        self._setup_data(R, M, K)
        self._setup_hyper(ARD, hyperparameters)

    def train(self, init_UV, iterations):
        ''' Initialise and run the sampler. '''
        self.initialise(init_UV=init_UV)
        self.run(iterations)

    def _draw_tau(self, alpha, beta):
        return gamma_draw(alpha, beta)

    def initialise(self, init_UV='random'):
        ''' Initialise U, V, tau, and lambda (if ARD).  (bnmf_gibbs.py:110-132) '''
        assert init_UV in OPTIONS_INIT_UV, "Unknown initialisation option: %s. Should be in %s." % (init_UV, OPTIONS_INIT_UV)

        self.lambdak = numpy.zeros(self.K)
        if self.ARD:
            for k in range(self.K):
                self.lambdak[k] = self.alpha0 / self.beta0

        eng = self._engine()
        self._upload_hyper(eng)
        eng.set_scalars(0., self.lambdak if self.ARD else None)
        eng.init_factors(self.METHOD, init_UV == 'random', next_seed() if init_UV == 'random' else 0)
        self.U = eng.get_factor(SIDE_U, F_VALUE)
        self.V = eng.get_factor(SIDE_V, F_VALUE)

        # Initialise tau
        self.tau = self._draw_tau(self.alpha_s(), self.beta_s())

    def _upload_state(self):
        eng = self._engine()
        self._upload_hyper(eng)
        eng.set_factor(SIDE_U, F_VALUE, self.U)
        eng.set_factor(SIDE_V, F_VALUE, self.V)
        eng.set_scalars(self.tau if hasattr(self, 'tau') else 0., self.lambdak if self.ARD else None)
        return eng

    _WANT_LAMBDAK = True

    def _run_begin(self, iterations):
        ''' Run the Gibbs sampler (bnmf_gibbs.py:135-183): reset the traces, upload the state. '''
        self.all_times = []  # to plot performance against time
        self.all_performances = {}  # for plotting convergence of metrics
        for metric in ALL_METRICS:
            self.all_performances[metric] = []

        eng = self._upload_state()
        seed = next_seed() if self.METHOD == _lib.GIBBS else 0
        traces = self._configure_summary(eng)
        return eng, seed, traces

    def _run_end(self, iterations, eng, res):
        traces = res['all_U'] is not None
        self._summary_ready = self._summary is not None and eng.summary_count() > 0

        self.all_U, self.all_V = res['all_U'], res['all_V']
        self.all_tau = res['stats'][:, _lib.ST_TAU].copy()
        if res['all_lambdak'] is None:      # batched run: the lambdak draws are not traced
            self.all_lambdak = None
        else:
            self.all_lambdak = res['all_lambdak'] if self.ARD else numpy.zeros((iterations, self.K))
        if iterations > 0:
            if traces:
                self.U, self.V = numpy.copy(self.all_U[-1]), numpy.copy(self.all_V[-1])
            else:
                self.U, self.V = eng.get_factor(SIDE_U, F_VALUE), eng.get_factor(SIDE_V, F_VALUE)
            self.tau = float(self.all_tau[-1])
            if self.ARD:
                self.lambdak = numpy.copy(self.all_lambdak[-1]) if self.all_lambdak is not None else eng.get_scalars()[1]
        self._fill_performances(eng, res['stats'])
        self.all_times = list(res['seconds'])
        if self.verbose:
            for it in range(iterations):
                print("Iteration %s. MSE: %s. R^2: %s. Rp: %s." % (
                    it + 1, self.all_performances['MSE'][it], self.all_performances['R^2'][it], self.all_performances['Rp'][it]))

    ''' Compute the parameters for the distributions we sample from. '''
    def alpha_s(self):
        ''' alpha* for tau. '''
        return self.alphatau + self.size_Omega / 2.0

    def beta_s(self):
        ''' beta* for tau. '''
        eng = self._upload_state()
        return self.betatau + 0.5 * eng.stats(self.METHOD)[_lib.ST_RSS]

    def alphak_s(self, k):
        ''' alphak* for lambdak. '''
        return self.alpha0 + self.I + self.J

    def betak_s(self, k):
        ''' betak* for lambdak. '''
        return self.beta0 + self.U[:, k].sum() + self.V[:, k].sum()

    def tauU(self, k):
        ''' tauUk for Uk. '''
        return self._upload_state().column_params(self.METHOD, SIDE_U, k)[0]

    def muU(self, tauUk, k):
        ''' muUk for Uk. '''
        return self._upload_state().column_params(self.METHOD, SIDE_U, k)[1]

    def tauV(self, k):
        ''' tauVk for Vk. '''
        return self._upload_state().column_params(self.METHOD, SIDE_V, k)[0]

    def muV(self, tauVk, k):
        ''' muVk for Vk. '''
        return self._upload_state().column_params(self.METHOD, SIDE_V, k)[1]

    def approx_expectation(self, burn_in, thinning):
        ''' Return our expectation of U, V, tau, lambdak. '''
        if self._summary_matches(burn_in, thinning):
            eng = self._engine()
            return (eng.get_summary(_lib.SUM_FACTOR_U), eng.get_summary(_lib.SUM_FACTOR_V),
                    float(eng.get_summary(_lib.SUM_TAU)[0]), eng.get_summary(_lib.SUM_LAMBDA_U) if self.ARD else None)
        self._need_traces(self.all_U)
        indices = range(burn_in, len(self.all_U), thinning)
        exp_U = numpy.array([self.all_U[i] for i in indices]).sum(axis=0) / float(len(indices))
        exp_V = numpy.array([self.all_V[i] for i in indices]).sum(axis=0) / float(len(indices))
        exp_tau = sum([self.all_tau[i] for i in indices]) / float(len(indices))
        exp_lambdak = None if not self.ARD else sum(
            [self.all_lambdak[i] for i in indices]) / float(len(indices))
        return (exp_U, exp_V, exp_tau, exp_lambdak)

    def predict(self, M_pred, burn_in, thinning):
        ''' Compute the expectation of U and V, and use it to predict missing values. '''
        if self._summary_matches(burn_in, thinning):
            return self._device_predict(M_pred, _lib.PRED_SUMMARY)
        (exp_U, exp_V, _, _) = self.approx_expectation(burn_in, thinning)
        return self._predict_from(M_pred, exp_U, exp_V)

    def predict_while_running(self):
        ''' Predict the training error while running. '''
        eng = self._upload_state()
        return eng.performances(eng.stats(self.METHOD))

    def quality(self, metric, burn_in, thinning):
        ''' Return the model quality, either as log likelihood, BIC, AIC, MSE, or ELBO. '''
        assert metric in ALL_QUALITY, 'Unrecognised metric for model quality: %s.' % metric

        if self._summary_matches(burn_in, thinning):
            # residual sum of squares of the summary means on Omega, evaluated on the device
            train = self._device_predict(self.M, _lib.PRED_SUMMARY)
            if metric == 'MSE':
                return train['MSE']
            exp_tau = float(self._engine().get_summary(_lib.SUM_TAU)[0])
            log_likelihood = self.size_Omega / 2. * (math.log(exp_tau) - math.log(2 * math.pi)) \
                - exp_tau / 2. * self._last_predict['SS_res']
        else:
            (exp_U, exp_V, exp_tau, _) = self.approx_expectation(burn_in, thinning)
            log_likelihood = self.log_likelihood(exp_U, exp_V, exp_tau)

        if metric == 'loglikelihood':
            return log_likelihood
        elif metric == 'BIC':
            # -2*loglikelihood + (no. free parameters * log(no data points))
            return - 2 * log_likelihood + self.number_parameters() * math.log(self.size_Omega)
        elif metric == 'AIC':
            # -2*loglikelihood + 2*no. free parameters
            return - 2 * log_likelihood + 2 * self.number_parameters()
        elif metric == 'MSE':
            R_pred = numpy.dot(exp_U, exp_V.T)
            return self.compute_MSE(self.M, self.R, R_pred)
        elif metric == 'ELBO':
            return 0.

    def log_likelihood(self, exp_U, exp_V, exp_tau):
        ''' Return the likelihood of the data given the trained model's parameters. '''
        exp_logtau = math.log(exp_tau)
        return self.size_Omega / 2. * (exp_logtau - math.log(2 * math.pi)) \
            - exp_tau / 2. * (self.M * (self.R - numpy.dot(exp_U, exp_V.T))**2).sum()

    def number_parameters(self):
        ''' Return the number of free variables in the model. '''
        return (self.I * self.K + self.J * self.K + 1) + (self.K if self.ARD else 0)
