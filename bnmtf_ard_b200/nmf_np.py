"""
Non-probabilistic non-negative matrix factorisation (Lee & Seung I-divergence
multiplicative updates) -- B200 engine.  Drop-in for the reference class
code/models/nmf_np.py: constructor (R, M, K), initialise(init_UV, expo_prior),
run(iterations), train(iterations, init_UV, expo_prior), predict(M_pred),
update_U(k), update_V(k), compute_I_div().
"""
import numpy

from . import _lib
from ._lib import F_VALUE, SIDE_U, SIDE_V
from ._nmf_common import ALL_METRICS, NMFCommon
from .distributions._device import exp_draws

OPTIONS_INIT_UV = ['ones', 'random', 'exponential']


class nmf_np(NMFCommon):
    METHOD = _lib.NP

    def __init__(self, R, M, K):
        ''' Set up the class and do some checks on the values passed. '''
        self.metrics = ['MSE', 'R^2', 'Rp']
        self._setup_data(R, M, K)
        # For computing the I-div it is easier if unknown values are 1's, not 0's, to avoid numerical issues
        # (nmf_np.py:54-57, vectorised)
        self.R_excl_unknown = numpy.where(self.M != 0, self.R, 1.)

    def train(self, iterations, init_UV='random', expo_prior=1.):
        ''' Initialise and run the algorithm. '''
        self.initialise(init_UV=init_UV, expo_prior=expo_prior)
        self.run(iterations=iterations)

    def initialise(self, init_UV='random', expo_prior=1.):
        ''' Initialise U and V.  (nmf_np.py:78-94) '''
        assert init_UV in OPTIONS_INIT_UV, "Unrecognised init option for U,V: %s. Should be one in %s." % (init_UV, OPTIONS_INIT_UV)

        if init_UV == 'ones':
            self.U = numpy.ones((self.I, self.K))
            self.V = numpy.ones((self.J, self.K))
        elif init_UV == 'random':
            self.U = numpy.random.rand(self.I, self.K)
            self.V = numpy.random.rand(self.J, self.K)
        elif init_UV == 'exponential':
            self.U = exp_draws(expo_prior * numpy.ones(self.I * self.K)).reshape(self.I, self.K)
            self.V = exp_draws(expo_prior * numpy.ones(self.J * self.K)).reshape(self.J, self.K)

    def _upload_state(self):
        eng = self._engine()
        eng.set_factor(SIDE_U, F_VALUE, self.U)
        eng.set_factor(SIDE_V, F_VALUE, self.V)
        return eng

    _WANT_LAMBDAK = False

    def _run_begin(self, iterations):
        ''' Run the algorithm (nmf_np.py:97-117): reset the traces, upload the state. '''
        assert hasattr(self, 'U') and hasattr(self, 'V'), "U and V have not been initialised - please run NMF.initialise() first."

        self.all_times = []  # to plot performance against time
        self.all_performances = {}  # for plotting convergence of metrics
        for metric in ALL_METRICS:
            self.all_performances[metric] = []
        return self._upload_state(), 0, False

    def _run_end(self, iterations, eng, res):
        self.U, self.V = eng.get_factor(SIDE_U, F_VALUE), eng.get_factor(SIDE_V, F_VALUE)
        self._fill_performances(eng, res['stats'])
        self.all_times = list(res['seconds'])
        self.all_i_div = [float(v) for v in res['stats'][:, _lib.ST_IDIV]]
        if self.verbose:
            for it in range(iterations):
                print("Iteration %s. I-divergence: %s. MSE: %s. R^2: %s. Rp: %s." % (
                    it + 1, self.all_i_div[it], self.all_performances['MSE'][it], self.all_performances['R^2'][it],
                    self.all_performances['Rp'][it]))

    ''' Updates for U and V. '''
    def update_U(self, k):
        ''' Update values for U.  (nmf_np.py:120-122) '''
        self.U[:, k] = self._upload_state().column_params(self.METHOD, SIDE_U, k)[1]

    def update_V(self, k):
        ''' Update values for V.  (nmf_np.py:124-126) '''
        self.V[:, k] = self._upload_state().column_params(self.METHOD, SIDE_V, k)[1]

    def predict(self, M_pred):
        ''' Predict missing values in R. '''
        if not self._shapes_fit_handle(U=(self.U, (self.I, self.K)), V=(self.V, (self.J, self.K))):
            return self._predict_from(M_pred, numpy.asarray(self.U, dtype=float), numpy.asarray(self.V, dtype=float))
        return self._device_predict(M_pred, _lib.PRED_CURRENT, self._upload_state())

    def compute_I_div(self):
        ''' Return the I-divergence.  (nmf_np.py:160-163) '''
        return float(self._upload_state().stats(self.METHOD)[_lib.ST_IDIV])

    def give_update(self, iteration):
        ''' Print and store the I-divergence and performances. '''
        perf = self.predict(self.M)
        i_div = self.compute_I_div()

        for metric in ALL_METRICS:
            self.all_performances[metric].append(perf[metric])

        print("Iteration %s. I-divergence: %s. MSE: %s. R^2: %s. Rp: %s." % (iteration, i_div, perf['MSE'], perf['R^2'], perf['Rp']))
