"""
Cross-validation with the fits of ALL parameter sets and folds running concurrently in
one batch on the GPU (BASELINE configs[4]).

Same constructor (plus an optional batch size), run(), find_best_parameters(), attributes
and log-file format as the reference's code/cross_validation/matrix_cross_validation.py.
The reference fits one model per (parameter set, fold) in turn (:83-101), or P at a time in
a multiprocessing.Pool (parallel_matrix_cross_validation.py:53-66).  Here the models are
constructed and initialised in exactly that order -- the folds come from the same `random`
stream and the engine seeds from the same numpy stream, so every fit equals the serial
driver's bit for bit -- but their device loops run together: one kernel launch per update
serves the whole batch (bnmtf_batch_run), replayed as one CUDA graph.

Models that are not classes of this package (no deferred run) are simply fitted in place.
"""
import time

from .. import _batch
from ._device import prepare_summary
from .matrix_cross_validation import MatrixCrossValidation


class BatchedMatrixCrossValidation(MatrixCrossValidation):
    def __init__(self, method, R, M, K, parameter_search, train_config, predict_config, file_performance,
                 batch_models=1024):
        MatrixCrossValidation.__init__(self, method, R, M, K, parameter_search, train_config, predict_config, file_performance)
        self.batch_models = int(batch_models)   # at most this many fits share a batch (device memory: ~24 B per observed entry and fit)
        self.launches = 0
        self.timing = {'setup_s': 0., 'device_loop_s': 0., 'predict_s': 0.}   # wall clock per phase

    def run(self):
        ''' Run the cross-validation. '''
        pending = []     # (parameters, [(model or None, test, performance or None)] or exception)
        n_models = 0
        floor = _batch.layout_floor([self.M])   # the training masks are subsets of M
        for parameters in self.parameter_search:
            t0 = time.perf_counter()
            if self.verbose:
                print("Trying parameters %s." % (parameters))
            try:
                folds_training, folds_test = self._folds()
                entries = []
                for i, (train, test) in enumerate(zip(folds_training, folds_test)):
                    if self.verbose:
                        print("Fold %s (parameters: %s)." % (i + 1, parameters))
                    model = self.method(self.R, train, **parameters)
                    prepare_summary(model, self.predict_config)
                    if _batch.supports_batching(model):
                        with _batch.deferred(model, floor):
                            model.train(**self.train_config)
                        entries.append((model, test, None))
                        n_models += 1
                    else:
                        model.train(**self.train_config)
                        entries.append((None, test, model.predict(test, **self.predict_config)))
                pending.append((parameters, entries))
            except Exception as e:
                pending.append((parameters, e))
            self.timing['setup_s'] += time.perf_counter() - t0
            if n_models >= self.batch_models:
                self._flush(pending)
                pending, n_models = [], 0
        self._flush(pending)

    def _flush(self, pending):
        ''' Run the deferred device loops of the pending fits together, then predict and log
            parameter set by parameter set, in the order of the serial driver. '''
        models = [m for (_, entries) in pending if not isinstance(entries, Exception) for (m, _, _) in entries if m is not None]
        failure = None
        t0 = time.perf_counter()
        if models:
            try:
                self.launches += _batch.run_deferred(models)
            except Exception as e:
                failure = e
        t1 = time.perf_counter()
        self.timing['device_loop_s'] += t1 - t0
        for parameters, entries in pending:
            try:
                if isinstance(entries, Exception):
                    raise entries
                if failure is not None and any(m is not None for (m, _, _) in entries):
                    raise failure
                self.all_performances[self.JSON(parameters)] = {}
                for (model, test, performance_dict) in entries:
                    if model is not None:
                        performance_dict = model.predict(test, **self.predict_config)
                    self.store_performances(performance_dict, parameters)
                self.log(parameters)
            except Exception as e:
                self.fout.write("Tried parameters %s but got exception: %s. \n" % (parameters, e))
                self.fout.flush()
        self.timing['predict_s'] += time.perf_counter() - t1
