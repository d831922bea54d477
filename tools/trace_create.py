"""Where the time of bnmf_create + the first calls goes (BNMF_TRACE laps), for a C3 shard of G genomes (dev tool)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["BNMF_TRACE"] = "1"
import numpy as np
from bayesnmf_b200 import Handle
from bayesnmf_b200.hyperpriors import fill_hyperprior_params
from tests.util import synth_counts
G = int(sys.argv[1]) if len(sys.argv) > 1 else 12500
M, _, _ = synth_counts(96, G, 20, 4000.0, seed=0)
M = np.asfortranarray(M)
for rep in range(3):
    t0 = time.time()
    h = Handle(M, 20, likelihood="poisson", prior="gamma", MH=False, seed=1, g0=0, G_total=100000)
    t1 = time.time()
    for k, v in fill_hyperprior_params(None, "gamma", float(M.mean()), 20).items():
        h.set_hyper(k, v)
    t2 = time.time()
    h.init_from_prior()
    t3 = time.time()
    h.step(5, want_P=True, want_A=True)
    t4 = time.time()
    print(f"rep {rep}: create {1e3*(t1-t0):.2f} ms, set_hyper x8 {1e3*(t2-t1):.2f} ms, init {1e3*(t3-t2):.2f} ms, 5 steps {1e3*(t4-t3):.2f} ms", file=sys.stderr)
    h.close()
