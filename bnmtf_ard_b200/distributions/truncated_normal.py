"""Scalar truncated normal on [0, inf) (reference:
code/models/distributions/truncated_normal.py:37-71); evaluated on the device."""
from ._device import tn_draw, tn_moments


def TN_draw(mu, tau):
    """tau == 0 -> 0.0; negative / non-finite draws -> 0.0 (truncated_normal.py:37-44)."""
    if tau == 0.:
        return 0.
    return float(tn_draw([mu], [tau])[0])


def TN_expectation(mu, tau):
    return float(tn_moments([mu], [tau])[0][0])


def TN_variance(mu, tau):
    return float(tn_moments([mu], [tau])[1][0])


def TN_mode(mu):
    return max(0.0, mu)
