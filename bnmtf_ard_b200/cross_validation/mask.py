"""
Methods for generating mask matrices, with 1 entries indicating observed and 0
indicating unobserved; single train/test splits and cross-validation folds.

Vectorised mirror of the reference's code/cross_validation/mask.py: same function
names, arguments, return values and assertion messages.  The reference walks all
I*J cells in Python (`nonzero_indices`, mask.py:21-24) and fills the masks entry by
entry -- minutes at 4*10^8 cells; here the index lists are numpy arrays.  The
random splits consume Python's `random` stream exactly like the reference
(`random.shuffle` of a list of the same length), so `random.seed(s)` reproduces
the reference's folds; above SHUFFLE_EXACT_LIMIT entries a numpy permutation
seeded from that stream is used instead (same distribution, different stream).
"""
import random

import numpy

SHUFFLE_EXACT_LIMIT = 5000000


''' Helpers. '''
def compute_Ms(folds_M):
    ''' Take in the ten fold M's, and construct the masks M for the other nine folds. '''
    no_folds = len(folds_M)
    folds_M = [numpy.array(fold_M) for fold_M in folds_M]
    return [sum(folds_M[:fold] + folds_M[fold + 1:]) for fold in range(0, no_folds)]


def nonzero_indices(M):
    ''' Return a list of indices of all nonzero indices in M (row-major order, like the reference). '''
    rows, cols = numpy.nonzero(numpy.array(M))
    return list(zip(rows.tolist(), cols.tolist()))


def _nonzero_flat(M):
    return numpy.flatnonzero(numpy.asarray(M))


def _shuffled(n):
    ''' The permutation random.shuffle applies to a list of length n (same stream as the reference). '''
    if n <= SHUFFLE_EXACT_LIMIT:
        perm = list(range(n))
        random.shuffle(perm)
        return numpy.array(perm, dtype=numpy.int64)
    return numpy.random.RandomState(random.getrandbits(32)).permutation(n)


def check_empty_rows_columns(M):
    ''' Return True if all rows and columns have at least one observation. '''
    M = numpy.asarray(M)
    return bool((M.sum(axis=1) != 0).all() and (M.sum(axis=0) != 0).all())


def calc_inverse_M(M, M_combined=None):
    ''' Return the inverse of M. If M_combined is defined, make sure the
        original plus inverse add up to be M_combined.'''
    return (M_combined if M_combined is not None else numpy.ones(M.shape)) - M


def _mask_from_flat(shape, flat):
    out = numpy.zeros(shape)
    out.flat[flat] = 1
    return out


''' Generating methods. '''
def generate_M(I, J, fraction, M=None):
    ''' Generate a mask matrix M_train with :fraction missing entries.
        If :M is defined, use only those 1-entries. '''
    if M is None:
        M = numpy.ones((I, J))
    flat = _nonzero_flat(M)
    no_elements = len(flat)
    no_missing_total = I * J - no_elements
    assert no_missing_total < I * J * fraction, "Specified %s fraction missing, so %s entries missing, but there are already %s missing by default!" % \
        (fraction, I * J * fraction, no_missing_total)

    # Shuffle the observed entries, take the first (I*J)*(1-fraction) and mark those as observed
    flat = flat[_shuffled(no_elements)]
    index_last_observed = int(I * J * (1 - fraction))
    M_train = _mask_from_flat((I, J), flat[:index_last_observed])
    M_test = _mask_from_flat((I, J), flat[index_last_observed:])
    assert numpy.array_equal(M, M_train + M_test), "Tried splitting M into M_test and M_train but something went wrong."
    return M_train, M_test


def try_generate_M(I, J, fraction, attempts, M=None):
    ''' Try generate_M() :attempts times, making sure each row and column has
        at least one observed entry. '''
    for i in range(attempts):
        M_train, M_test = generate_M(I=I, J=J, fraction=fraction, M=M)
        if check_empty_rows_columns(M_train):
            return M_train, M_test
    assert False, "Failed to generate folds for training and test data, %s attempts, fraction %s." % (attempts, fraction)


def compute_folds(I, J, no_folds, M=None):
    ''' Compute :no_folds cross-validation masks.
        Return a tuple (Ms_train, Ms_test), both a list of masks.
        If M is defined, split only the 1 entries into the folds. '''
    if M is None:
        M = numpy.ones((I, J))
    M = numpy.asarray(M)

    no_elements = M.sum()
    flat = _nonzero_flat(M)
    flat = flat[_shuffled(len(flat))]
    split_places = [int(i * no_elements / no_folds) for i in range(0, no_folds + 1)]  # find the indices where the next fold start

    Ms_train, Ms_test = [], []  # list of the M's for the different folds
    for i in range(0, no_folds):
        M_test = _mask_from_flat(M.shape, flat[split_places[i]:split_places[i + 1]])
        Ms_train.append(M - M_test)
        Ms_test.append(M_test)
    return (Ms_train, Ms_test)


def _attempts(fold_function, I, J, no_folds, attempts, M):
    for i in range(attempts):
        Ms_train, Ms_test = fold_function(I=I, J=J, no_folds=no_folds, M=M)
        if all(check_empty_rows_columns(M_train) for M_train in Ms_train):
            return Ms_train, Ms_test
    assert False, "Failed to generate folds for training and test data, %s attempts." % attempts


def compute_folds_attempts(I, J, no_folds, attempts, M=None):
    ''' Try compute_folds() :attempts times, making sure each row and column of
        M_train has at least one observed entry. '''
    return _attempts(compute_folds, I, J, no_folds, attempts, M)


def compute_folds_stratify_rows(I, J, no_folds, M=None):
    ''' Run compute_folds() but make sure each fold has the same number of
        entries of each row - i.e. we stratify the folds based on rows.
        This could still create an empty column though. '''
    if M is None:
        M = numpy.ones((I, J))
    M = numpy.asarray(M)

    # Stratified folds with the row index as the label (the reference calls scikit-learn's
    # StratifiedKFold(labels, n_folds, shuffle=True) here, mask.py:89-104)
    M = numpy.ascontiguousarray(M, dtype=float)
    flat = _nonzero_flat(M)
    labels = flat // M.shape[1]
    test_folds = _stratified_test_folds(labels, no_folds)
    fold_of = numpy.full(M.size, -1, dtype=numpy.int8 if no_folds < 128 else numpy.int32)
    fold_of[flat] = test_folds
    fold_of = fold_of.reshape(M.shape)
    Ms_train, Ms_test = [], []
    for fold in range(no_folds):
        M_test = (fold_of == fold).astype(float)
        Ms_train.append(M - M_test), Ms_test.append(M_test)
    return Ms_train, Ms_test


def _stratified_test_folds(labels, n_splits):
    ''' Fold index of every sample, exactly as scikit-learn's StratifiedKFold(n_splits,
        shuffle=True) with random_state=None assigns it (same draws from numpy's global
        RandomState: one shuffle per class, classes in order of appearance) -- for labels that
        are sorted, which lets each class be a slice instead of a mask over all samples
        (scikit-learn's loop is O(samples x classes): 0.5 s per call at 887 x 545). '''
    labels = numpy.asarray(labels)
    n = labels.size
    assert n == 0 or bool(numpy.all(labels[1:] >= labels[:-1])), "labels must be sorted"
    starts = numpy.flatnonzero(numpy.concatenate(([True], labels[1:] != labels[:-1]))) if n else numpy.zeros(0, dtype=int)
    ends = numpy.concatenate((starts[1:], [n])).astype(int)
    test_folds = numpy.empty(n, dtype="i")
    splits = numpy.arange(n_splits)
    for a, b in zip(starts, ends):
        # samples a..b-1 of the sorted order fall into folds p % n_splits (round robin over the sorted labels)
        first = a % n_splits
        counts = numpy.bincount((first + numpy.arange(b - a)) % n_splits, minlength=n_splits)
        folds_for_class = splits.repeat(counts)
        numpy.random.shuffle(folds_for_class)
        test_folds[a:b] = folds_for_class
    return test_folds


def compute_folds_stratify_rows_attempts(I, J, no_folds, attempts, M=None):
    ''' Try compute_folds_stratify_rows() :attempts times, making sure each row
        and column of M_train has at least one observed entry. '''
    return _attempts(compute_folds_stratify_rows, I, J, no_folds, attempts, M)


def compute_folds_stratify_columns(I, J, no_folds, M=None):
    ''' Same as compute_folds_stratify_rows() but now stratify by column. '''
    Ms_train_T, Ms_test_T = compute_folds_stratify_rows(I=J, J=I, no_folds=no_folds, M=M.T if M is not None else None)
    Ms_train, Ms_test = [M_train.T for M_train in Ms_train_T], [M_test.T for M_test in Ms_test_T]
    return (Ms_train, Ms_test)


def compute_folds_stratify_columns_attempts(I, J, no_folds, attempts, M=None):
    ''' Try compute_folds_stratify_columns() :attempts times, making sure each row
        and column of M_train has at least one observed entry. '''
    return _attempts(compute_folds_stratify_columns, I, J, no_folds, attempts, M)
