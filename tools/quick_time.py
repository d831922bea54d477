"""Quick device timing of the Poisson-Gamma iteration and of k_zstat alone (dev tool)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bayesnmf_b200 import Handle
from tests.util import synth_counts

def run(K, G, N, mu, iters=20, prior="gamma"):
    M, P, E = synth_counts(K, G, N, mu, seed=0)
    t0 = time.time()
    h = Handle(M, N, likelihood="poisson", prior=prior, MH=False, seed=1, ring_cap=0)
    t1 = time.time()
    from oracle.gibbs import default_hyperprior_params
    for k, v in default_hyperprior_params(prior, M.mean(), N).items():
        h.set_hyper(k[0].upper() + k[1:], v)
    h.init_from_prior()
    h.step(5)
    zs = [h.sample_z(100 + i) for i in range(5)]
    h.step(3)
    t2 = time.time()
    out = h.step(iters)
    t3 = time.time()
    tm = h.timing()
    alg = 4.0 * (K * G + 2 * N * G + 2 * K * N)
    print(f"K={K} G={G} N={N} mu={mu}: create {t1-t0:.2f}s  z-only ms {np.round(zs,3)}  "
          f"step {tm['total_ms']/iters:.3f} ms/iter (zstat {tm['zstat_ms']/iters:.3f}) wall {1e3*(t3-t2)/iters:.3f} "
          f"launches/iter {tm['launches']/iters:.1f}  picks {M.sum():.3g}  z roofline frac {alg/1e9/(min(zs)*1e-3)/6545.3:.4f} "
          f"RMSE {out['metrics'][-1][1]:.3f}", flush=True)

if __name__ == "__main__":
    run(96, 100, 5, 4000.0, iters=200)
    run(96, 100000, 20, 4000.0)
    run(96, 100000, 20, 100.0)
    run(96, 20000, 15, 4000.0)
