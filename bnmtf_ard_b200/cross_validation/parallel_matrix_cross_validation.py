"""
Cross-validation with the folds of one parameter set running concurrently (mirror of
the reference's code/cross_validation/parallel_matrix_cross_validation.py).

The reference forks P worker processes (multiprocessing.Pool, :53-66), which cannot
be done once a CUDA context exists.  Here the P workers are threads: every fold's
model owns its device handle and CUDA stream and the ctypes calls release the GIL,
so the small kernels of different folds overlap on the GPU.  Like the reference's
run_model (:28-31) the prediction is called without predict_config.
"""
from concurrent.futures import ThreadPoolExecutor

import numpy

from .matrix_cross_validation import MatrixCrossValidation


def run_fold(params):
    (parameters, R, train, test, method, train_config, predict_config) = \
        (params['parameters'], params['R'], params['train'], params['test'], params['method'], params['train_config'], params['predict_config'])
    performance_dict = run_model(method, R, train, test, parameters, train_config, predict_config)
    return performance_dict


# Method for running the model with the given parameters
def run_model(method, X, train, test, parameters, train_config, predict_config):
    model = method(X, train, **parameters)
    model.train(**train_config)
    return model.predict(test)


# Class, redefining the run function
class ParallelMatrixCrossValidation(MatrixCrossValidation):
    def __init__(self, method, R, M, K, parameter_search, train_config, predict_config, file_performance, P):
        MatrixCrossValidation.__init__(self, method, R, M, K, parameter_search, train_config, predict_config, file_performance)
        self.P = P

    # Run the cross-validation
    def run(self):
        for parameters in self.parameter_search:
            if self.verbose:
                print("Trying parameters %s." % (parameters))

            try:
                folds_training, folds_test = self._folds()

                # We need to put the parameter dict into json to hash it
                self.all_performances[self.JSON(parameters)] = {}

                # Create the workers for the folds, and run them
                all_parameters = [
                    {
                        'parameters': parameters,
                        'R': self.R,
                        'train': train,
                        'test': test,
                        'method': self.method,
                        'train_config': self.train_config,
                        'predict_config': self.predict_config,
                    }
                    for (train, test) in zip(folds_training, folds_test)
                ]
                with ThreadPoolExecutor(max_workers=self.P) as pool:
                    outputs = list(pool.map(run_fold, all_parameters))

                for performance_dict in outputs:
                    self.store_performances(performance_dict, parameters)

                self.log(parameters)

            except Exception as e:
                self.fout.write("Tried parameters %s but got exception: %s. \n" % (parameters, e))

    # Undo the function run_model:
    def run_model(self, train, test, parameters):
        raise Exception("Using wrong method for ParallelMatrixCrossValidation! Use the one defined outside of the class.")
