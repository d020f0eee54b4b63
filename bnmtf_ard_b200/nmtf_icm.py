"""
Iterated conditional modes (MAP) for non-negative matrix tri-factorisation, with
ARD -- B200 engine.  Drop-in for the reference class code/models/nmtf_icm.py: the
BNMTF Gibbs loop with every draw replaced by the mode of its conditional and a
floor of MINIMUM_TN on F, S and G (nmtf_icm.py:186-217).
"""
from . import _lib
from .bnmtf_gibbs import bnmtf_gibbs
from .distributions.gamma import gamma_mode

OPTIONS_INIT_FG = ['kmeans', 'random', 'exp']
OPTIONS_INIT_S = ['random', 'exp']
MINIMUM_TN = 0.1


class nmtf_icm(bnmtf_gibbs):
    METHOD = _lib.ICM

    def _draw_tau(self, alpha, beta):
        return gamma_mode(alpha, beta)
