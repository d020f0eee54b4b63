"""
General framework for performing cross validation for a matrix prediction method
(mirror of the reference's code/cross_validation/matrix_cross_validation.py: same
constructor, run(), find_best_parameters(), attributes and log-file format).

- method: a class with constructor (R, M, **parameters), train(**train_config) and
  predict(M_test, **predict_config) -> {'MSE','R^2','Rp'}; the classes of
  bnmtf_ard_b200 (or the reference's own) fit.
"""
import json

import numpy

from .mask import compute_folds_stratify_columns_attempts, compute_folds_stratify_rows_attempts
from ._device import prepare_summary

attempts_generate_M = 1000


class MatrixCrossValidation:
    verbose = True

    def __init__(self, method, R, M, K, parameter_search, train_config, predict_config, file_performance):
        self.method = method
        self.R = numpy.array(R, dtype=float)
        self.M = numpy.array(M)
        self.K = K
        self.train_config = train_config
        self.predict_config = predict_config
        self.parameter_search = parameter_search

        self.fout = open(file_performance, 'w')
        (self.I, self.J) = self.R.shape
        assert (self.R.shape == self.M.shape), "X and M are of different shapes: %s and %s respectively." % (self.R.shape, self.M.shape)

        self.all_performances = {}      # Performances across all folds - mapping JSON of parameters to a dictionary from evaluation criteria to a list of performances
        self.average_performances = {}  # Average performances across folds - mapping JSON of parameters to a dictionary from evaluation criteria to average performance
        self.performances = {}          # Average performances per criterion - mapping evaluation criterion to a list of average performances (same size as parameter_search)

    def _folds(self):
        folds_method = compute_folds_stratify_rows_attempts if self.I < self.J else compute_folds_stratify_columns_attempts
        return folds_method(I=self.I, J=self.J, no_folds=self.K, attempts=attempts_generate_M, M=self.M)

    def run(self):
        ''' Run the cross-validation. '''
        for parameters in self.parameter_search:
            if self.verbose:
                print("Trying parameters %s." % (parameters))

            try:
                folds_training, folds_test = self._folds()

                # We need to put the parameter dict into json to hash it
                self.all_performances[self.JSON(parameters)] = {}
                for i, (train, test) in enumerate(zip(folds_training, folds_test)):
                    if self.verbose:
                        print("Fold %s (parameters: %s)." % (i + 1, parameters))
                    performance_dict = self.run_model(train, test, parameters)
                    self.store_performances(performance_dict, parameters)

                self.log(parameters)

            except Exception as e:
                self.fout.write("Tried parameters %s but got exception: %s. \n" % (parameters, e))
                self.fout.flush()

    def run_model(self, train, test, parameters):
        ''' Initialises and runs the model, and returns the performance on the test set. '''
        model = self.method(self.R, train, **parameters)
        prepare_summary(model, self.predict_config)
        model.train(**self.train_config)
        return model.predict(test, **self.predict_config)

    def JSON(self, d):
        ''' Returns the sorted json of the dictionary given. '''
        # Cannot handle numpy arrays so force all numpy arrays to be lists
        d_copy = d.copy()
        for key, val in d.items():
            if isinstance(val, numpy.ndarray):
                d_copy[key] = val.tolist()
        return json.dumps(d_copy, sort_keys=True)

    def store_performances(self, performance_dict, parameters):
        ''' Store the performances we get back in a dictionary from criterion name to a list of performances. '''
        for name in performance_dict:
            if name in self.all_performances[self.JSON(parameters)]:
                self.all_performances[self.JSON(parameters)][name].append(performance_dict[name])
            else:
                self.all_performances[self.JSON(parameters)][name] = [performance_dict[name]]

    def compute_average_performances(self, parameters):
        ''' Compute the average performance of the given parameters, across the K folds. '''
        performances = self.all_performances[self.JSON(parameters)]
        average_performances = {name: (sum(values) / float(len(values))) for (name, values) in performances.items()}
        self.average_performances[self.JSON(parameters)] = average_performances

        # Also store a dictionary from evaluation criterion to a list of average performances
        for (name, avr_perf) in average_performances.items():
            if name in self.performances:
                self.performances[name].append(avr_perf)
            else:
                self.performances[name] = [avr_perf]

    def find_best_parameters(self, evaluation_criterion, low_better):
        ''' Finds the parameter values of the best performance for the specified criterion. '''
        min_or_max = min if low_better else max
        self.best_performance = min_or_max(self.performances[evaluation_criterion])
        index_best = self.performances[evaluation_criterion].index(self.best_performance)

        self.best_parameters = self.parameter_search[index_best]
        self.best_performances_all = self.average_performances[self.JSON(self.best_parameters)]

        self.log_best(index_best)
        return (self.best_parameters, self.best_performance)

    def log(self, parameters):
        ''' Logs the performances for specific parameter values. '''
        self.compute_average_performances(parameters)
        message = "Tried parameters %s. Average performances: %s. \nAll performances: %s. \n" % (
            parameters, self.average_performances[self.JSON(parameters)], self.all_performances[self.JSON(parameters)])
        self.fout.write(message)
        self.fout.flush()

    def log_best(self, index_best):
        ''' Logs the best performances and parameters. '''
        message = "Best performances: %s. Best parameters: %s. \n" % (self.best_performances_all, self.best_parameters)
        self.fout.write(message)
        self.fout.flush()
