"""Second-handle construction cost after a long run of a first handle (dev tool)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests.util import synth_counts
from bayesnmf_b200 import Handle
M, _, _ = synth_counts(96, 100000, 20, 4000.0, seed=0)
M = np.asfortranarray(M, dtype=np.float64)
steps = int(sys.argv[1]); flush = int(sys.argv[2])
h = Handle(M, 20, likelihood="poisson", prior="gamma", MH=False, seed=1)
h.init_from_prior()
if flush: h.set_l2_flush(512 << 20)
h.step(steps)
if flush: h.set_l2_flush(0)
h.step(steps)
print("--- second handle", flush=True)
t0 = time.time()
h2 = Handle(M, 20, likelihood="poisson", prior="gamma", MH=False, seed=1)
print(f"Handle() {1e3*(time.time()-t0):.1f} ms", flush=True)
