"""
Nested cross-validation (mirror of the reference's
code/cross_validation/nested_matrix_cross_validation.py): for each outer fold an inner
(parallel or serial) cross-validation picks the best parameters by MSE, then the model is
trained on the outer training set and evaluated on the outer test set.
"""
import numpy

from .mask import compute_folds_stratify_columns_attempts, compute_folds_stratify_rows_attempts
from .matrix_cross_validation import MatrixCrossValidation
from .parallel_matrix_cross_validation import ParallelMatrixCrossValidation
from .batched_matrix_cross_validation import BatchedMatrixCrossValidation
from ._device import prepare_summary

attempts_generate_M = 1000


class MatrixNestedCrossValidation:
    verbose = True

    def __init__(self, method, R, M, K, P, parameter_search, train_config, predict_config, file_performance, files_nested_performances):
        self.method = method
        self.R = numpy.array(R, dtype=float)
        self.M = numpy.array(M)
        self.K = K
        self.P = P
        self.train_config = train_config
        self.predict_config = predict_config
        self.parameter_search = parameter_search
        self.files_nested_performances = files_nested_performances

        self.fout = open(file_performance, 'w')
        (self.I, self.J) = self.R.shape
        assert (self.R.shape == self.M.shape), "X and M are of different shapes: %s and %s respectively." % (self.R.shape, self.M.shape)

        self.all_performances = {}      # Performances across all folds - dictionary from evaluation criteria to a list of performances
        self.average_performances = {}  # Average performances across folds - dictionary from evaluation criteria to average performance

    def run(self, parallel=True):
        ''' Run the cross-validation. '''
        folds_method = compute_folds_stratify_rows_attempts if self.I < self.J else compute_folds_stratify_columns_attempts
        folds_training, folds_test = folds_method(I=self.I, J=self.J, no_folds=self.K, attempts=attempts_generate_M, M=self.M)

        for i, (train, test) in enumerate(zip(folds_training, folds_test)):
            if self.verbose:
                print("Fold %s of nested cross-validation." % (i + 1))

            # Run the cross-validation (parallel='batched': all inner fits of the fold share one device loop)
            crossval = BatchedMatrixCrossValidation(
                method=self.method,
                R=self.R,
                M=train,
                K=self.K,
                parameter_search=self.parameter_search,
                train_config=self.train_config,
                predict_config=self.predict_config,
                file_performance=self.files_nested_performances[i],
            ) if parallel == 'batched' else ParallelMatrixCrossValidation(
                method=self.method,
                R=self.R,
                M=train,
                K=self.K,
                parameter_search=self.parameter_search,
                train_config=self.train_config,
                predict_config=self.predict_config,
                file_performance=self.files_nested_performances[i],
                P=self.P,
            ) if parallel else MatrixCrossValidation(
                method=self.method,
                R=self.R,
                M=train,
                K=self.K,
                parameter_search=self.parameter_search,
                train_config=self.train_config,
                predict_config=self.predict_config,
                file_performance=self.files_nested_performances[i],
            )
            crossval.verbose = self.verbose
            crossval.run()

            try:
                (best_parameters, _) = crossval.find_best_parameters(evaluation_criterion='MSE', low_better=True)
                if self.verbose:
                    print("Best parameters for fold %s were %s." % (i + 1, best_parameters))
            except KeyError:
                best_parameters = self.parameter_search[0]
                if self.verbose:
                    print("Found no performances, dataset too sparse? Use first values instead for fold %s, %s." % (i + 1, best_parameters))

            # Train the model and test the performance on the test set
            performance_dict = self.run_model(train, test, best_parameters)
            self.store_performances(performance_dict)
            if self.verbose:
                print("Finished fold %s, with performances %s." % (i + 1, performance_dict))

        self.log()

    def run_model(self, train, test, parameters):
        ''' Initialises and runs the model, and returns the performance on the test set. '''
        model = self.method(self.R, train, **parameters)
        prepare_summary(model, self.predict_config)
        model.train(**self.train_config)
        return model.predict(test, **self.predict_config)

    def store_performances(self, performance_dict):
        ''' Store the performances we get back in a dictionary from criterion name to a list of performances. '''
        for name in performance_dict:
            if name in self.all_performances:
                self.all_performances[name].append(performance_dict[name])
            else:
                self.all_performances[name] = [performance_dict[name]]

    def compute_average_performances(self):
        ''' Compute the average performance of the given parameters, across the K folds. '''
        performances = self.all_performances
        average_performances = {name: (sum(values) / float(len(values))) for (name, values) in performances.items()}
        self.average_performances = average_performances

    def log(self):
        ''' Logs the performance on the test set for this fold. '''
        self.compute_average_performances()
        message = "Average performances: %s. \nAll performances: %s. \n" % (self.average_performances, self.all_performances)
        self.fout.write(message)
        self.fout.flush()
