"""Normal draw (reference: code/models/distributions/normal.py:8-10).  Not on the
hot path of any model (no reference model imports it); kept for API parity and
drawn from numpy's global stream exactly as the reference does."""
import math

import numpy


def normal_draw(mu, tau):
    sigma = numpy.float64(1.0) / math.sqrt(tau)
    return numpy.random.normal(loc=mu, scale=sigma, size=None)
