"""How far the drop-in is from the reference's exact `==` literals in the few reference tests that compare floats for
equality (run on the GPU box; prints relative differences).  Uses the loader of tests/test_gpu_reference_suite.py."""
import os
import sys

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import test_gpu_reference_suite as rs  # noqa: E402


def rel(a, b):
    a, b = numpy.asarray(a, dtype=float), numpy.asarray(b, dtype=float)
    return float(numpy.max(numpy.abs(a - b) / numpy.maximum(numpy.abs(b), 1e-300)))


m = rs._load("test_bnmf_vb.py")
g = m.__dict__
BNMF = m.bnmf_vb(g["R"], g["M"], g["K"], False, g["hyperparams"])
BNMF.exp_U, BNMF.exp_V = 1. / g["lambdaU"], 1. / g["lambdaV"]
BNMF.var_U, BNMF.var_V = numpy.ones((g["I"], g["K"])) * 2, numpy.ones((g["J"], g["K"])) * 3
print("bnmf_vb exp_square_diff: ours %r, reference literal %r, rel %.3g" % (BNMF.exp_square_diff(), 172.66666666666666, rel(BNMF.exp_square_diff(), 172.66666666666666)))
BNMF.mu_U, BNMF.tau_U = numpy.zeros((g["I"], g["K"])), numpy.zeros((g["I"], g["K"]))
BNMF.exp_tau = 3.
k = 0
BNMF.update_U(k)
M, R = g["M"], g["R"]
want_tau = numpy.array([3. * (M[i] * (BNMF.exp_V[:, k] * BNMF.exp_V[:, k] + BNMF.var_V[:, k])).sum() for i in range(g["I"])])
want_mu = numpy.array([(1. / want_tau[i]) * (-2. + BNMF.exp_tau * (M[i] * ((BNMF.R[i] - numpy.dot(BNMF.exp_U[i], BNMF.exp_V.T) + BNMF.exp_U[i, k] * BNMF.exp_V[:, k]) * BNMF.exp_V[:, k])).sum()) for i in range(g["I"])])
print("bnmf_vb update_U: tau rel %.3g, mu rel %.3g" % (rel(BNMF.tau_U[:, k], want_tau), rel(BNMF.mu_U[:, k], want_mu)))

m = rs._load("test_nmtf_np.py")
import math
R = numpy.array([[1, 2], [3, 4]], dtype=float); M = numpy.array([[1, 1], [0, 1]])
nm = m.nmtf_np(R, M, 3, 1)
nm.F, nm.S, nm.G = numpy.array([[1, 2, 3], [4, 5, 6]]), numpy.array([[7], [8], [9]]), numpy.array([[10], [11]])
want = sum([1.0 * math.log(1.0 / 500.0) - 1.0 + 500.0, 2.0 * math.log(2.0 / 550.0) - 2.0 + 550.0, 4.0 * math.log(4.0 / 1342.0) - 4.0 + 1342.0])
print("nmtf_np compute_I_div: ours %r, reference %r, rel %.3g" % (nm.compute_I_div(), want, rel(nm.compute_I_div(), want)))
