"""Vector truncated normal on [0, inf) (reference:
code/models/distributions/truncated_normal_vector.py:37-78); one device launch
per call instead of one Python call per element."""
import numpy

from ._device import tn_draw, tn_moments


def TN_vector_draw(mus, taus):
    return list(tn_draw(mus, taus))


def TN_vector_expectation(mus, taus):
    return list(tn_moments(mus, taus)[0])


def TN_vector_variance(mus, taus):
    return list(tn_moments(mus, taus)[1])


def TN_vector_mode(mus):
    zeros = numpy.zeros(len(mus))
    return numpy.maximum(zeros, mus)
