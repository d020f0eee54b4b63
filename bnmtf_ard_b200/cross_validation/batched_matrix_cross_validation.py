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
import threading
import time
from concurrent.futures import ThreadPoolExecutor

from .. import _batch
from ._device import prepare_summary
from .matrix_cross_validation import MatrixCrossValidation


class BatchedMatrixCrossValidation(MatrixCrossValidation):
    def __init__(self, method, R, M, K, parameter_search, train_config, predict_config, file_performance,
                 batch_models=128, workers=8):
        MatrixCrossValidation.__init__(self, method, R, M, K, parameter_search, train_config, predict_config, file_performance)
        self.batch_models = int(batch_models)   # a batch is run once it holds this many fits; while its device loop runs, the next batch is set up
        self.workers = max(1, int(workers))     # host threads that pack the models' data and evaluate their predictions
        self.launches = 0
        self.device_seconds = 0.     # CUDA-event time of the batched device loops
        self.timing = {'setup_s': 0., 'device_loop_s': 0., 'predict_s': 0.}   # wall clock per phase

    def run(self):
        ''' Run the cross-validation. '''
        pending = []     # (parameters, [(model or None, test, performance or None)] or exception)
        n_models = 0
        floor = _batch.layout_floor([self.M])   # the training masks are subsets of M
        self._flusher = None
        try:   # one growth step of the device memory pool for the handles of two batches in flight
            n_fits = min(2 * self.batch_models + self.K, len(self.parameter_search) * self.K)
            _batch.reserve_device_memory(n_fits, self.I, self.J, int((self.M != 0).sum()), self.workers)
        except Exception:
            pass    # no device: the models raise their own error
        for parameters in self.parameter_search:
            t0 = time.perf_counter()
            if self.verbose:
                print("Trying parameters %s." % (parameters))
            try:
                folds_training, folds_test = self._folds()
                entries = []
                # constructors (copies and checks of R, M) and device handles (upload and packing) do not touch the
                # random streams: build them concurrently
                def build(train):
                    m = self.method(self.R, train, **parameters)
                    if _batch.supports_batching(m):
                        m._layout_floor = floor
                        m._engine()
                    return m
                models = list(self._pool().map(build, folds_training)) if self.workers > 1 else [build(t) for t in folds_training]
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
                self._submit(pending)
                pending, n_models = [], 0
        self._submit(pending)
        self._join()
        if self._executor is not None:
            self._executor.shutdown()
            self._executor = None

    _executor = None
    _flusher = None

    def _submit(self, pending):
        ''' Run a batch in the background (its device loop, downloads, predictions and log lines) while the
            caller sets up the next one.  One batch in flight: the log keeps the order of the parameter sets. '''
        self._join()
        if not pending:
            return
        self._flusher = threading.Thread(target=self._flush, args=(pending,))
        self._flusher.start()

    def _join(self):
        if self._flusher is not None:
            self._flusher.join()
            self._flusher = None

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
                self.device_seconds += sum(_batch.last_device_seconds)
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
