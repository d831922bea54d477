"""Device timing of the BASELINE.json configurations other than C3 (dev tool)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bayesnmf_b200 import Handle
from bayesnmf_b200.sampler import get_temp_sched
from tests.util import synth_counts

def run(name, K, G, N, lik, prior, MH, learn=False, iters=20, mu=4000.0, prec="f64"):
    M, P, E = synth_counts(K, G, N if not learn else max(2, N // 2), mu, seed=0)
    if lik == "normal":
        M = M + np.random.default_rng(1).normal(0, 0.05 * M.mean(axis=0, keepdims=True) + 1e-3, M.shape)
    h = Handle(M, N, likelihood=lik, prior=prior, MH=MH, learning_rank=learn, seed=1, precision=prec)
    if learn:
        h.set_temperature_schedule(get_temp_sched(5000, 1000, np.random.default_rng(0)))
    h.init_from_prior()
    h.step(3)
    out = h.step(iters); t = h.timing()
    msg = f"{name}: K={K} G={G} N={N} {lik}-{prior} MH={MH} learn={learn}: {t['iter_ms']/iters:.3f} ms/iter ({1e3*iters/t['iter_ms']:.0f} it/s), {t['launches']/iters:.0f} launches/iter"
    if MH:
        h.step(2, converged=True)
        out = h.step(iters, converged=True); t = h.timing()
        msg += f" | real MH step: {t['iter_ms']/iters:.3f} ms/iter, {t['launches']/iters:.0f} launches/iter, P acc {out['metrics'][-1][9]:.3f}"
    print(msg, "RMSE", round(out["metrics"][-1][1], 3), flush=True)
    h.close()

if __name__ == "__main__":
    which = sys.argv[1:] or ["c1", "c2", "c4", "c5"]
    if "c1" in which: run("c1", 96, 100, 5, "poisson", "gamma", False, iters=200)
    if "c2" in which: run("c2", 96, 500, 10, "poisson", "truncnormal", True, learn=True, iters=100)
    if "c4" in which: run("c4", 96, 20000, 15, "normal", "truncnormal", False, iters=30)
    if "c5" in which: run("c5", 1536, 50000, 40, "poisson", "exponential", True, iters=5)
