"""
Non-probabilistic non-negative matrix tri-factorisation (Yoo & Choi I-divergence
multiplicative updates) -- B200 engine.  Drop-in for the reference class
code/models/nmtf_np.py: constructor (R, M, K, L), initialise(init_FG, init_S,
expo_prior), run(iterations), train(iterations, init_FG, init_S, expo_prior),
predict(M_pred), update_F(k), update_S(k,l), update_G(l), compute_I_div().
"""
import numpy

from . import _lib
from ._engine import NMTFEngine
from ._lib import F_VALUE, SIDE_U, SIDE_V
from ._nmf_common import ALL_METRICS, NMFCommon, check_nmtf_limits
from .distributions._device import exp_draws

OPTIONS_INIT_FG = ['kmeans', 'ones', 'random', 'exponential']
OPTIONS_INIT_S = ['ones', 'random', 'exponential']


class nmtf_np(NMFCommon):
    METHOD = _lib.NP

    def __init__(self, R, M, K, L):
        ''' Set up the class and do some checks on the values passed.  (nmtf_np.py:40-63) '''
        self.metrics = ['MSE', 'R^2', 'Rp']
        self._setup_data(R, M, K)
        self.L = L
        check_nmtf_limits(K, L)
        # For computing the I-div it is easier if unknown values are 1's, not 0's, to avoid numerical issues
        self.R_excl_unknown = numpy.where(self.M != 0, self.R, 1.)

    def _engine(self):
        if self._eng is None:
            self._eng = NMTFEngine(self.R, self.M, self.K, self.L)
        return self._eng

    def train(self, iterations, init_FG='random', init_S='random', expo_prior=1.):
        ''' Initialise and run the algorithm. '''
        self.initialise(init_FG=init_FG, init_S=init_S, expo_prior=expo_prior)
        self.run(iterations=iterations)

    def initialise(self, init_FG='random', init_S='random', expo_prior=1.):
        ''' Initialise F, S and G.  (nmtf_np.py:83-123) '''
        assert init_FG in OPTIONS_INIT_FG, "Unrecognised init option for F,G: %s. Should be one in %s." % (init_FG, OPTIONS_INIT_FG)
        assert init_S in OPTIONS_INIT_S, "Unrecognised init option for S: %s. Should be one in %s." % (init_S, OPTIONS_INIT_S)

        if init_S == 'ones':
            self.S = numpy.ones((self.K, self.L))
        elif init_S == 'random':
            self.S = numpy.random.rand(self.K, self.L)
        elif init_S == 'exponential':
            self.S = exp_draws(expo_prior * numpy.ones(self.K * self.L)).reshape(self.K, self.L)

        if init_FG == 'ones':
            self.F = numpy.ones((self.I, self.K))
            self.G = numpy.ones((self.J, self.L))
        elif init_FG == 'random':
            self.F = numpy.random.rand(self.I, self.K)
            self.G = numpy.random.rand(self.J, self.L)
        elif init_FG == 'exponential':
            self.F = exp_draws(expo_prior * numpy.ones(self.I * self.K)).reshape(self.I, self.K)
            self.G = exp_draws(expo_prior * numpy.ones(self.J * self.L)).reshape(self.J, self.L)
        elif init_FG == 'kmeans':
            # nmtf_np.py:113-123: one-hot cluster indicators + 0.2
            from .kmeans.kmeans import KMeans
            eng = self._engine()
            kmeans_F = KMeans(self.R, self.M, self.K, _engine=eng, _side=SIDE_U)
            kmeans_F.initialise()
            kmeans_F.cluster()
            self.F = kmeans_F.clustering_results + 0.2
            kmeans_G = KMeans(self.R.T, self.M.T, self.L, _engine=eng, _side=SIDE_V)
            kmeans_G.initialise()
            kmeans_G.cluster()
            self.G = kmeans_G.clustering_results + 0.2

    def _upload_state(self):
        eng = self._engine()
        eng.set_factor(SIDE_U, F_VALUE, self.F)
        eng.set_factor(SIDE_V, F_VALUE, self.G)
        eng.set_S(F_VALUE, self.S)
        return eng

    def run(self, iterations):
        ''' Run the algorithm.  (nmtf_np.py:126-160) '''
        assert hasattr(self, 'F') and hasattr(self, 'S') and hasattr(self, 'G'), \
            "F, S and G have not been initialised - please run NMTF.initialise() first."

        self.all_times = []  # to plot performance against time
        self.all_performances = {}  # for plotting convergence of metrics
        for metric in ALL_METRICS:
            self.all_performances[metric] = []

        eng = self._upload_state()
        res = eng.np_run(iterations)
        self.F, self.G, self.S = eng.get_factor(SIDE_U, F_VALUE), eng.get_factor(SIDE_V, F_VALUE), eng.get_S(F_VALUE)
        self._fill_performances(eng, res['stats'])
        self.all_times = list(res['seconds'])
        self.all_i_div = [float(v) for v in res['stats'][:, _lib.ST_IDIV]]
        if self.verbose:
            for it in range(iterations):
                print("Iteration %s. I-divergence: %s. MSE: %s. R^2: %s. Rp: %s." % (
                    it + 1, self.all_i_div[it], self.all_performances['MSE'][it], self.all_performances['R^2'][it],
                    self.all_performances['Rp'][it]))

    ''' Updates for F, G, S. '''
    def triple_dot(self, M1, M2, M3):
        ''' Triple matrix multiplication: M1*M2*M3.  (nmtf_np.py:163-171) '''
        K, L = M2.shape
        if K < L:
            return numpy.dot(M1, numpy.dot(M2, M3))
        else:
            return numpy.dot(numpy.dot(M1, M2), M3)

    def update_F(self, k):
        ''' Update values for F.  (nmtf_np.py:173-179) '''
        self.F[:, k] = self._upload_state().np_update(0, k=k)

    def update_G(self, l):
        ''' Update values for G.  (nmtf_np.py:181-187) '''
        self.G[:, l] = self._upload_state().np_update(1, l=l)

    def update_S(self, k, l):
        ''' Update values for S.  (nmtf_np.py:189-195) '''
        self.S[k, l] = float(self._upload_state().np_update(2, k=k, l=l)[0])

    def predict(self, M_pred):
        ''' Predict missing values in R. '''
        if not self._shapes_fit_handle(F=(self.F, (self.I, self.K)), S=(self.S, (self.K, self.L)), G=(self.G, (self.J, self.L))):
            FS = numpy.dot(numpy.asarray(self.F, dtype=float), numpy.asarray(self.S, dtype=float))
            return self._predict_from(M_pred, FS, numpy.asarray(self.G, dtype=float))
        return self._device_predict(M_pred, _lib.PRED_CURRENT, self._upload_state())

    def compute_I_div(self):
        ''' Return the I-divergence.  (nmtf_np.py:222-225) '''
        return float(self._upload_state().stats(self.METHOD)[_lib.ST_IDIV])

    def give_update(self, iteration):
        ''' Print and store the I-divergence and performances. '''
        perf = self.predict(self.M)
        i_div = self.compute_I_div()
        for metric in ALL_METRICS:
            self.all_performances[metric].append(perf[metric])
        print("Iteration %s. I-divergence: %s. MSE: %s. R^2: %s. Rp: %s." % (iteration, i_div, perf['MSE'], perf['R^2'], perf['Rp']))
