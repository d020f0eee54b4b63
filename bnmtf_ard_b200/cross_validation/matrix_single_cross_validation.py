"""
Cross-validation of ONE parameter setting (mirror of the reference's
code/cross_validation/matrix_single_cross_validation.py): K stratified folds, the
performance of every fold and their average are written to file_performance.
"""
import numpy

from .mask import compute_folds_stratify_columns_attempts, compute_folds_stratify_rows_attempts
from ._device import prepare_summary

ATTEMPTS_GENERATE_M = 1000
METRICS = ['MSE', 'R^2', 'Rp']


class MatrixSingleCrossValidation:
    verbose = True

    def __init__(self, method, R, M, K, parameters, train_config, predict_config, file_performance):
        self.method = method
        self.R = numpy.array(R, dtype=float)
        self.M = numpy.array(M)
        self.K = K
        self.parameters = parameters
        self.train_config = train_config
        self.predict_config = predict_config

        self.fout = open(file_performance, 'w')
        (self.I, self.J) = self.R.shape
        assert (self.R.shape == self.M.shape), "X and M are of different shapes: %s and %s respectively." % (self.R.shape, self.M.shape)

        # Performances across all folds - dictionary from evaluation criteria to a list of performances
        self.performances = {metric: [] for metric in METRICS}

    def run(self):
        ''' Run the cross-validation. '''
        if self.verbose:
            print("Running cross-validation framework.")

        # Compute the mask matrices
        folds_method = compute_folds_stratify_rows_attempts if self.I < self.J else compute_folds_stratify_columns_attempts
        folds_training, folds_test = folds_method(I=self.I, J=self.J, no_folds=self.K, attempts=ATTEMPTS_GENERATE_M, M=self.M)

        # Run each fold and store the performances.
        for i, (train, test) in enumerate(zip(folds_training, folds_test)):
            if self.verbose:
                print("Fold %s." % (i + 1))
            performance_dict = self.run_model(train, test, self.parameters)
            self.log_performance(i + 1, performance_dict)
        self.log_average_performance()

    def run_model(self, train, test, parameters):
        ''' Initialises and runs the model, and returns the performance on the test set. '''
        model = self.method(self.R, train, **parameters)
        prepare_summary(model, self.predict_config)
        model.train(**self.train_config)
        return model.predict(test, **self.predict_config)

    def log_performance(self, fold, performance_dict):
        ''' Logs the best performances and parameters. '''
        message = "Performances fold %s: %s. \n" % (fold, performance_dict)
        self.fout.write(message)
        self.fout.flush()
        for metric in METRICS:
            self.performances[metric].append(performance_dict[metric])

    def log_average_performance(self):
        avr_performance = {metric: numpy.mean(self.performances[metric]) for metric in METRICS}
        message = "Average performance: %s. All performances: %s." % (avr_performance, self.performances)
        self.fout.write(message)
        self.fout.flush()
