"""Where the fixed cost of a run goes: construct + upload, prior draw, a few steps, final E (dev tool)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests.util import synth_counts
from bayesnmf_b200 import Handle
from bayesnmf_b200.hyperpriors import fill_hyperprior_params
M, _, _ = synth_counts(96, 100000, 20, 4000.0, seed=0)
M = np.asfortranarray(M, dtype=np.float64)
for rep in range(3):
    t = [time.time()]
    h = Handle(M, 20, likelihood="poisson", prior="gamma", MH=False, seed=1); t.append(time.time())
    mean = float(M.sum()) / M.size; t.append(time.time())
    for k, v in fill_hyperprior_params(None, "gamma", mean, 20).items():
        h.set_hyper(k, v)
    t.append(time.time())
    h.init_from_prior(); t.append(time.time())
    h.step(10, want_P=True, want_A=True); t.append(time.time())
    E = h.get_state("E"); t.append(time.time())
    h.close(); t.append(time.time())
    names = ["Handle()", "mean", "set_hyper x8", "init_from_prior", "10 steps", "get E", "close"]
    print(rep, " ".join(f"{n} {1e3*(b-a):.1f}ms" for n, a, b in zip(names, t[:-1], t[1:])), flush=True)
