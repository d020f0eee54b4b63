"""Gamma(shape alpha, rate beta): draws on the device, closed forms on the host
(reference: code/models/distributions/gamma.py:11-29)."""
import math

from scipy.special import psi as digamma

from ._device import gamma_draws


def gamma_draw(alpha, beta):
    return float(gamma_draws([float(alpha)], [float(beta)])[0])


def gamma_expectation(alpha, beta):
    alpha, beta = float(alpha), float(beta)
    return alpha / beta


def gamma_expectation_log(alpha, beta):
    alpha, beta = float(alpha), float(beta)
    return digamma(alpha) - math.log(beta)


def gamma_mode(alpha, beta):
    alpha, beta = float(alpha), float(beta)
    return (alpha - 1) / beta
