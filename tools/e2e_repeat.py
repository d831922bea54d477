"""The end-to-end span of bench.py (construct, prior draw, steps with every sample to the host, final E)
repeated in one process: its spread on a box (dev tool)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests.util import synth_counts
from bayesnmf_b200 import Handle
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
M, _, _ = synth_counts(96, 100000, 20, 4000.0, seed=0)
M = np.asfortranarray(M, dtype=np.float64)
for r in range(reps):
    t0 = time.time()
    h = Handle(M, 20, likelihood="poisson", prior="gamma", MH=False, seed=1)
    ta = time.time()
    h.init_from_prior()
    o = h.step(steps, want_P=True, want_A=True)
    tb = time.time()
    E = h.get_state("E")
    t1 = time.time()
    h.close()
    if len(sys.argv) > 3 and r == 2:
        from bayesnmf_b200 import release_cached_memory
        release_cached_memory()
    print(f"rep {r}: construct {1e3*(ta-t0):.1f} ms, prior + {steps} steps {1e3*(tb-ta):.1f} ms, final E {1e3*(t1-tb):.1f} ms, "
          f"total {1e3*(t1-t0):.1f} ms = {steps/(t1-t0):.0f} it/s", flush=True)
