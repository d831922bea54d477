"""Small runs of every kernel family (all priors, MH, rank learning, ring + MAP) -- dev tool; the pool does not allow compute-sanitizer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bayesnmf_b200 import Handle
from bayesnmf_b200.sampler import get_temp_sched
from tests.util import synth_counts

def run(K, G, N, lik, prior, MH, learn=False, mu=300.0):
    M, _, _ = synth_counts(K, G, N, mu, seed=0)
    if lik == "normal":
        M = M + np.random.default_rng(1).normal(0, 1.0, M.shape)
    h = Handle(M, N, likelihood=lik, prior=prior, MH=MH, learning_rank=learn, seed=1, ring_cap=4)
    if learn:
        h.set_temperature_schedule(get_temp_sched(50, 10, np.random.default_rng(0)))
    h.init_from_prior()
    h.step(3)
    if MH:
        h.step(2, converged=True)
    h.get_map(3)
    print("ok", K, G, N, lik, prior, MH, learn, flush=True)
    h.close()

run(96, 130, 5, "poisson", "gamma", False)
run(40, 70, 20, "poisson", "exponential", False, learn=True)
run(96, 300, 4, "poisson", "truncnormal", True, learn=True)
run(10, 9000, 3, "poisson", "exponential", True)          # a row over several blocks of a cluster
run(96, 5000, 6, "normal", "truncnormal", False)          # Gram path, 8 lanes per genome
run(200, 64, 3, "normal", "exponential", False)
