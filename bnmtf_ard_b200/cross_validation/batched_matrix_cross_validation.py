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
from concurrent.futures import ThreadPoolExecutor

from .. import _batch
from ._device import prepare_summary
from .matrix_cross_validation import MatrixCrossValidation


class BatchedMatrixCrossValidation(MatrixCrossValidation):
    def __init__(self, method, R, M, K, parameter_search, train_config, predict_config, file_performance,
                 batch_models=1024, workers=8):
        MatrixCrossValidation.__init__(self, method, R, M, K, parameter_search, train_config, predict_config, file_performance)
        self.batch_models = int(batch_models)   # at most this many fits share a batch (device memory: ~24 B per observed entry and fit)
        self.workers = max(1, int(workers))     # host threads that pack the models' data and evaluate their predictions
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
                models = [self.method(self.R, train, **parameters) for train in folds_training]
                batchable = [m for m in models if _batch.supports_batching(m)]
                for m in batchable:
                    m._layout_floor = floor
                # device handles (upload and packing of R, M) are independent of the random streams: build them concurrently
                if len(batchable) > 1 and self.workers > 1:
                    list(self._pool().map(lambda m: m._engine(), batchable))
                # initialise() and the seed of run() consume numpy's global stream: serial, in the order of the serial driver
                for i, (model, test) in enumerate(zip(models, folds_test)):
                    if self.verbose:
                        print("Fold %s (parameters: %s)." % (i + 1, parameters))
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
        if self._executor is not None:
            self._executor.shutdown()
            self._executor = None

    _executor = None

    def _pool(self):
        if self._executor is None:
            self._executor = ThreadPoolExecutor(max_workers=self.workers)
        return self._executor

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
        # held-out predictions: independent per model, evaluated on the device from `workers` host threads
        jobs = [(model, test) for (_, entries) in pending if not isinstance(entries, Exception)
                for (model, test, _) in entries if model is not None] if failure is None else []

        def predict(job):
            try:
                return job[0].predict(job[1], **self.predict_config)
            except Exception as e:      # reported with the parameter set it belongs to
                return e
        predicted = {}
        if jobs:
            results = list(self._pool().map(predict, jobs)) if self.workers > 1 else [predict(j) for j in jobs]
            predicted = {id(job[0]): res for job, res in zip(jobs, results)}
        for parameters, entries in pending:
            try:
                if isinstance(entries, Exception):
                    raise entries
                if failure is not None and any(m is not None for (m, _, _) in entries):
                    raise failure
                self.all_performances[self.JSON(parameters)] = {}
                for (model, test, performance_dict) in entries:
                    if model is not None:
                        performance_dict = predicted[id(model)]
                        if isinstance(performance_dict, Exception):
                            raise performance_dict
                    self.store_performances(performance_dict, parameters)
                self.log(parameters)
            except Exception as e:
                self.fout.write("Tried parameters %s but got exception: %s. \n" % (parameters, e))
                self.fout.flush()
        self.timing['predict_s'] += time.perf_counter() - t1
