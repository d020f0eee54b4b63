"""Exponential draws (reference: code/models/distributions/exponential.py:7-9)."""
from ._device import exp_draws


def exponential_draw(lambdax):
    """One draw of Exp(rate lambdax) from the device Philox stream."""
    return float(exp_draws([lambdax])[0])
