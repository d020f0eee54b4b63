"""Device-backed mirror of the reference's code/models/distributions package."""
from .exponential import exponential_draw  # noqa: F401
from .gamma import gamma_draw, gamma_expectation, gamma_expectation_log, gamma_mode  # noqa: F401
from .normal import normal_draw  # noqa: F401
from .truncated_normal import TN_draw, TN_expectation, TN_mode, TN_variance  # noqa: F401
from .truncated_normal_vector import (TN_vector_draw, TN_vector_expectation, TN_vector_mode,  # noqa: F401
                                      TN_vector_variance)
