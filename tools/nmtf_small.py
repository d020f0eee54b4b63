import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
from bnmtf_ard_b200 import bnmtf_gibbs, nmtf_icm
rng = np.random.default_rng(1)
for (I, J, K, L) in ((100, 80, 5, 5), (707, 139, 10, 10)):
    R = rng.exponential(1, (I, K)) @ rng.exponential(1, (K, L)) @ rng.exponential(1, (J, L)).T + rng.normal(0, 1, (I, J))
    M = (rng.random((I, J)) < 0.8).astype(float); M[:, 0] = 1; M[0, :] = 1
    hp = {'alphatau': 1., 'betatau': 1., 'alpha0': 1., 'beta0': 1., 'lambdaS': np.ones((K, L))}
    for cls in (bnmtf_gibbs, nmtf_icm):
        np.random.seed(0)
        m = cls(R, M, K, L, True, hp); m.verbose = False
        m.initialise('random', 'random'); m.run(20)
        m.run(500)
        print(cls.__name__, (I, J, K, L), "it/s %.0f" % (500 / m.all_times[-1]), "mse %.4f" % m.all_performances['MSE'][-1])
from bnmtf_ard_b200 import bnmtf_vb, nmtf_np
for (I, J, K, L) in ((100, 80, 5, 5), (707, 139, 10, 10)):
    R = np.abs(rng.exponential(1, (I, K)) @ rng.exponential(1, (K, L)) @ rng.exponential(1, (J, L)).T + rng.normal(0, 1, (I, J))) + 0.05
    M = (rng.random((I, J)) < 0.8).astype(float); M[:, 0] = 1; M[0, :] = 1
    hp = {'alphatau': 1., 'betatau': 1., 'alpha0': 1., 'beta0': 1., 'lambdaS': np.ones((K, L))}
    np.random.seed(0)
    m = bnmtf_vb(R, M, K, L, True, hp); m.verbose = False
    m.initialise('random', 'random'); m.run(10); m.run(200)
    print("bnmtf_vb", (I, J, K, L), "it/s %.0f" % (200 / m.all_times[-1]), "mse %.6f" % m.all_performances['MSE'][-1])
    np.random.seed(0)
    m = nmtf_np(R, M, K, L); m.verbose = False
    m.initialise('random', 'random'); m.run(10); m.run(200)
    print("nmtf_np", (I, J, K, L), "it/s %.0f" % (200 / m.all_times[-1]), "mse %.6f" % m.all_performances['MSE'][-1])
