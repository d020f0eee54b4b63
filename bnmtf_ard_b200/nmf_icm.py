"""
Iterated conditional modes (MAP) for non-negative matrix factorisation, with ARD
-- B200 engine.  Drop-in for the reference class code/models/nmf_icm.py: the
Gibbs loop with every draw replaced by the mode of its conditional and a floor
of MINIMUM_TN on the factors (nmf_icm.py:59,148-172).
"""
from . import _lib
from .bnmf_gibbs import bnmf_gibbs
from .distributions.gamma import gamma_mode

OPTIONS_INIT_UV = ['random', 'exp']
MINIMUM_TN = 0.1  # ICM has the tendency to set most columns to 0's; we reset them to this value.


class nmf_icm(bnmf_gibbs):
    METHOD = _lib.ICM

    def _draw_tau(self, alpha, beta):
        return gamma_mode(alpha, beta)
