"""B200-native engine for the BNMF / NMF models of ThomasBrouwer/BNMTF_ARD.

The classes mirror the reference's code/models API and run their inference
loops on the GPU through the C ABI in include/bnmtf_b200.h (ctypes, no torch
types); see DESIGN.md.  There is no CPU fallback.
"""
from .bnmf_gibbs import bnmf_gibbs  # noqa: F401
from .bnmf_vb import bnmf_vb  # noqa: F401
from .nmf_icm import nmf_icm  # noqa: F401
from .nmf_np import nmf_np  # noqa: F401
from .bnmtf_gibbs import bnmtf_gibbs  # noqa: F401
from .nmtf_icm import nmtf_icm  # noqa: F401
from .bnmtf_vb import bnmtf_vb  # noqa: F401
from .nmtf_np import nmtf_np  # noqa: F401

__all__ = ["bnmf_gibbs", "bnmf_vb", "nmf_icm", "nmf_np", "bnmtf_gibbs", "nmtf_icm", "bnmtf_vb", "nmtf_np"]
