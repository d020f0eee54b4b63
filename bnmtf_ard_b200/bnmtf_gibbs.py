"""
Gibbs sampler for non-negative matrix tri-factorisation (R ~ F S G^T), with ARD
-- B200 engine.  Drop-in for the reference class code/models/bnmtf_gibbs.py: same
constructor (R, M, K, L, ARD, hyperparameters), initialise(init_FG, init_S),
run(iterations), train(init_FG, init_S, iterations), predict(M_pred, burn_in,
thinning), quality(...), the public attributes (F, S, G, tau, lambdaFk, lambdaGl,
all_F, all_S, all_G, all_tau, all_lambdaFk, all_lambdaGl, all_times,
all_performances) and the per-parameter methods tauF/muF/tauS/muS/tauG/muG.
"""
import math

import numpy

from . import _lib
from ._engine import NMTFEngine
from ._lib import F_VALUE, SIDE_U, SIDE_V
from ._nmf_common import ALL_METRICS, ALL_QUALITY, NMFCommon, check_nmtf_limits
from .distributions._device import next_seed
from .distributions.gamma import gamma_draw, gamma_mode

OPTIONS_INIT_FG = ['kmeans', 'random', 'exp']
OPTIONS_INIT_S = ['random', 'exp']


class bnmtf_gibbs(NMFCommon):
    METHOD = _lib.GIBBS

    def __init__(self, R, M, K, L, ARD, hyperparameters):
        ''' Set up the class and do some checks on the values passed.  (bnmtf_gibbs.py:66-99) '''
        self._setup_data(R, M, K)
        self.L = L
        check_nmtf_limits(K, L)
        self.ARD = ARD

        self.alphatau, self.betatau = float(hyperparameters['alphatau']), float(hyperparameters['betatau'])
        self.lambdaS = numpy.array(hyperparameters['lambdaS'])
        if self.lambdaS.shape == ():
            self.lambdaS = self.lambdaS * numpy.ones((self.K, self.L))
        assert self.lambdaS.shape == (self.K, self.L), "Prior matrix lambdaS has the wrong shape: %s instead of (%s, %s)." % (self.lambdaS.shape, self.K, self.L)

        if self.ARD:
            self.alpha0, self.beta0 = float(hyperparameters['alpha0']), float(hyperparameters['beta0'])
        else:
            self.lambdaF, self.lambdaG = numpy.array(hyperparameters['lambdaF']), numpy.array(hyperparameters['lambdaG'])
            # Make lambdaF/G into a numpy array if they are a float
            if self.lambdaF.shape == ():
                self.lambdaF = self.lambdaF * numpy.ones((self.I, self.K))
            if self.lambdaG.shape == ():
                self.lambdaG = self.lambdaG * numpy.ones((self.J, self.L))

            assert self.lambdaF.shape == (self.I, self.K), "Prior matrix lambdaF has the wrong shape: %s instead of (%s, %s)." % (self.lambdaF.shape, self.I, self.K)
            assert self.lambdaG.shape == (self.J, self.L), "Prior matrix lambdaG has the wrong shape: %s instead of (%s, %s)." % (self.lambdaG.shape, self.J, self.L)

    # -- engine ----------------------------------------------------------------
    def _engine(self):
        if self._eng is None:
            self._eng = NMTFEngine(self.R, self.M, self.K, self.L)
        return self._eng

    def _upload_hyper(self, eng):
        if self.ARD:
            eng.set_hyper(True, self.alphatau, self.betatau, self.alpha0, self.beta0, None, None, self.lambdaS)
        else:
            eng.set_hyper(False, self.alphatau, self.betatau, 1., 1., self.lambdaF, self.lambdaG, self.lambdaS)

    def _upload_state(self):
        eng = self._engine()
        self._upload_hyper(eng)
        eng.set_factor(SIDE_U, F_VALUE, self.F)
        eng.set_factor(SIDE_V, F_VALUE, self.G)
        eng.set_S(F_VALUE, self.S)
        eng.set_scalars(self.tau if hasattr(self, 'tau') else 0.,
                        self.lambdaFk if self.ARD else None, self.lambdaGl if self.ARD else None)
        return eng

    def train(self, init_FG, init_S, iterations):
        ''' Initialise and run the sampler. '''
        self.initialise(init_FG=init_FG, init_S=init_S)
        self.run(iterations)

    def _draw_tau(self, alpha, beta):
        return gamma_draw(alpha, beta)

    def initialise(self, init_FG='random', init_S='random'):
        ''' Initialise F, S, G, tau, and lambdaFk, lambdaGl (if ARD).  (bnmtf_gibbs.py:121-167) '''
        assert init_FG in OPTIONS_INIT_FG, "Unknown initialisation option for F and G: %s. Should be in %s." % (init_FG, OPTIONS_INIT_FG)
        assert init_S in OPTIONS_INIT_S, "Unknown initialisation option for S: %s. Should be in %s." % (init_S, OPTIONS_INIT_S)
        self.lambdaFk = numpy.zeros(self.K)
        self.lambdaGl = numpy.zeros(self.L)
        if self.ARD:
            for k in range(self.K):
                self.lambdaFk[k] = self.alpha0 / self.beta0
            for l in range(self.L):
                self.lambdaGl[l] = self.alpha0 / self.beta0

        eng = self._engine()
        self._upload_hyper(eng)
        eng.set_scalars(0., self.lambdaFk if self.ARD else None, self.lambdaGl if self.ARD else None)
        eng.init(self.METHOD, init_FG == 'random', init_S == 'random',
                 next_seed() if init_FG == 'random' else 0, next_seed() if init_S == 'random' else 0)
        self.F = eng.get_factor(SIDE_U, F_VALUE)
        self.G = eng.get_factor(SIDE_V, F_VALUE)
        self.S = eng.get_S(F_VALUE)
        if init_FG == 'kmeans':
            # bnmtf_gibbs.py:140-151: one-hot cluster indicators + 0.2
            self.F, self.G = self._kmeans_FG(eng)
            self.F, self.G = self.F + 0.2, self.G + 0.2

        # Initialise tau
        self.tau = self._draw_tau(self.alpha_s(), self.beta_s())

    def _kmeans_FG(self, eng):
        ''' K-means clustering of the rows (K clusters) and of the columns (L clusters) of R on the device,
            over the engine's packed entries (code/models/kmeans/kmeans.py). '''
        from .kmeans.kmeans import KMeans
        if self.verbose:
            print("Initialising F using KMeans.")
        kmeans_F = KMeans(self.R, self.M, self.K, _engine=eng, _side=SIDE_U)
        kmeans_F.initialise()
        kmeans_F.cluster()
        if self.verbose:
            print("Initialising G using KMeans.")
        kmeans_G = KMeans(self.R.T, self.M.T, self.L, _engine=eng, _side=SIDE_V)
        kmeans_G.initialise()
        kmeans_G.cluster()
        return kmeans_F.clustering_results, kmeans_G.clustering_results

    def run(self, iterations):
        ''' Run the Gibbs sampler.  (bnmtf_gibbs.py:170-229) '''
        self.all_times = []  # to plot performance against time
        self.all_performances = {}  # for plotting convergence of metrics
        for metric in ALL_METRICS:
            self.all_performances[metric] = []

        eng = self._upload_state()
        seed = next_seed() if self.METHOD == _lib.GIBBS else 0
        traces = self._configure_summary(eng)
        res = eng.run(self.METHOD, iterations, seed=seed, traces=traces)
        self._summary_ready = self._summary is not None and eng.summary_count() > 0

        self.all_F, self.all_S, self.all_G = res['all_F'], res['all_S'], res['all_G']
        self.all_tau = res['stats'][:, _lib.ST_TAU].copy()
        self.all_lambdaFk = res['all_lambdaFk'] if self.ARD else numpy.zeros((iterations, self.K))
        self.all_lambdaGl = res['all_lambdaGl'] if self.ARD else numpy.zeros((iterations, self.L))
        if iterations > 0:
            if traces:
                self.F, self.S, self.G = numpy.copy(self.all_F[-1]), numpy.copy(self.all_S[-1]), numpy.copy(self.all_G[-1])
            else:
                self.F, self.G = eng.get_factor(0, _lib.F_VALUE), eng.get_factor(1, _lib.F_VALUE)
                self.S = eng.get_S(_lib.F_VALUE)
            self.tau = float(self.all_tau[-1])
            if self.ARD:
                self.lambdaFk, self.lambdaGl = numpy.copy(self.all_lambdaFk[-1]), numpy.copy(self.all_lambdaGl[-1])
        self._fill_performances(eng, res['stats'])
        self.all_times = list(res['seconds'])
        if self.verbose:
            for it in range(iterations):
                print("Iteration %s. MSE: %s. R^2: %s. Rp: %s." % (
                    it + 1, self.all_performances['MSE'][it], self.all_performances['R^2'][it], self.all_performances['Rp'][it]))

    def triple_dot(self, M1, M2, M3):
        ''' Triple matrix multiplication: M1*M2*M3.  (bnmtf_gibbs.py:232-240) '''
        K, L = M2.shape
        if K < L:
            return numpy.dot(M1, numpy.dot(M2, M3))
        else:
            return numpy.dot(numpy.dot(M1, M2), M3)

    ''' Compute the parameters for the distributions we sample from. '''
    def alpha_s(self):
        ''' alpha* for tau. '''
        return self.alphatau + self.size_Omega / 2.0

    def beta_s(self):
        ''' beta* for tau. '''
        eng = self._upload_state()
        return self.betatau + 0.5 * eng.stats(self.METHOD)[_lib.ST_RSS]

    def alphaFk_s(self, k):
        ''' alphaFk* for lambdaFk. '''
        return self.alpha0 + self.I

    def betaFk_s(self, k):
        ''' betak* for lambdaFk. '''
        return self.beta0 + self.F[:, k].sum()

    def alphaGl_s(self, l):
        ''' alphaFk* for lambdaFk. '''
        return self.alpha0 + self.J

    def betaGl_s(self, l):
        ''' betak* for lambdaFk. '''
        return self.beta0 + self.G[:, l].sum()

    def tauF(self, k):
        ''' tauFk for Fk. '''
        return self._upload_state().column_params(self.METHOD, 0, k=k)[0]

    def muF(self, tauFk, k):
        ''' muFk for Fk. '''
        return self._upload_state().column_params(self.METHOD, 0, k=k)[1]

    def tauS(self, k, l):
        ''' tauSkl for Skl. '''
        return float(self._upload_state().column_params(self.METHOD, 2, k=k, l=l)[0][0])

    def muS(self, tauSkl, k, l):
        ''' muSkl for Skl. '''
        return float(self._upload_state().column_params(self.METHOD, 2, k=k, l=l)[1][0])

    def tauG(self, l):
        ''' tauGl for Gl. '''
        return self._upload_state().column_params(self.METHOD, 1, l=l)[0]

    def muG(self, tauGl, l):
        ''' muGl for Gl. '''
        return self._upload_state().column_params(self.METHOD, 1, l=l)[1]

    def approx_expectation(self, burn_in, thinning):
        ''' Return our expectation of F, S, G, tau, lambdaFk, lambdaGl. '''
        if self._summary_matches(burn_in, thinning):
            eng = self._engine()
            return (eng.get_summary(_lib.SUM_FACTOR_U), eng.get_summary(_lib.SUM_S), eng.get_summary(_lib.SUM_FACTOR_V),
                    float(eng.get_summary(_lib.SUM_TAU)[0]),
                    eng.get_summary(_lib.SUM_LAMBDA_U) if self.ARD else None,
                    eng.get_summary(_lib.SUM_LAMBDA_V) if self.ARD else None)
        self._need_traces(self.all_F)
        indices = range(burn_in, len(self.all_F), thinning)
        exp_F = numpy.array([self.all_F[i] for i in indices]).sum(axis=0) / float(len(indices))
        exp_S = numpy.array([self.all_S[i] for i in indices]).sum(axis=0) / float(len(indices))
        exp_G = numpy.array([self.all_G[i] for i in indices]).sum(axis=0) / float(len(indices))
        exp_tau = sum([self.all_tau[i] for i in indices]) / float(len(indices))
        exp_lambdaFk = None if not self.ARD else sum([self.all_lambdaFk[i] for i in indices]) / float(len(indices))
        exp_lambdaGl = None if not self.ARD else sum([self.all_lambdaGl[i] for i in indices]) / float(len(indices))
        return (exp_F, exp_S, exp_G, exp_tau, exp_lambdaFk, exp_lambdaGl)

    def predict(self, M_pred, burn_in, thinning):
        ''' Compute the expectation of F, S, G, and use it to predict missing values. '''
        if self._summary_matches(burn_in, thinning):
            return self._device_predict(M_pred, _lib.PRED_SUMMARY)
        (exp_F, exp_S, exp_G, _, _, _) = self.approx_expectation(burn_in, thinning)
        R_pred = self.triple_dot(exp_F, exp_S, exp_G.T)
        MSE = self.compute_MSE(M_pred, self.R, R_pred)
        R2 = self.compute_R2(M_pred, self.R, R_pred)
        Rp = self.compute_Rp(M_pred, self.R, R_pred)
        return {'MSE': MSE, 'R^2': R2, 'Rp': Rp}

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
            (exp_F, exp_S, exp_G, exp_tau, _, _) = self.approx_expectation(burn_in, thinning)
            log_likelihood = self.log_likelihood(exp_F, exp_S, exp_G, exp_tau)

        if metric == 'loglikelihood':
            return log_likelihood
        elif metric == 'BIC':
            # -2*loglikelihood + (no. free parameters * log(no data points))
            return - 2 * log_likelihood + self.number_parameters() * math.log(self.size_Omega)
        elif metric == 'AIC':
            # -2*loglikelihood + 2*no. free parameters
            return - 2 * log_likelihood + 2 * self.number_parameters()
        elif metric == 'MSE':
            R_pred = self.triple_dot(exp_F, exp_S, exp_G.T)
            return self.compute_MSE(self.M, self.R, R_pred)
        elif metric == 'ELBO':
            return 0.

    def log_likelihood(self, exp_F, exp_S, exp_G, exp_tau):
        ''' Return the likelihood of the data given the trained model's parameters. '''
        exp_logtau = math.log(exp_tau)
        return self.size_Omega / 2. * (exp_logtau - math.log(2 * math.pi)) \
            - exp_tau / 2. * (self.M * (self.R - self.triple_dot(exp_F, exp_S, exp_G.T))**2).sum()

    def number_parameters(self):
        ''' Return the number of free variables in the model. '''
        return (self.I * self.K + self.K * self.L + self.J * self.L + 1) + (self.K + self.L if self.ARD else 0)
