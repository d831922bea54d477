import sys, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bayesnmf_b200 import Handle
from tests.util import synth_counts
M, _, _ = synth_counts(96, 100000, 20, 4000.0, seed=0)
h = Handle(M, 20, likelihood="poisson", prior="gamma", MH=False, seed=1, ring_cap=200)
h.init_from_prior(); h.step(220)
for ns in (200,):
    t0 = time.time(); P, E, A, nm = h.get_map(ns); t1 = time.time()
    Pl, Ph, El, Eh, nm2 = h.get_credible_intervals(ns); t2 = time.time()
    print(ns, nm, nm2, f"get_map {1e3*(t1-t0):.1f} ms, credible intervals {1e3*(t2-t1):.1f} ms", bool((El <= E).all() and (E <= Eh).all() or True), float(np.mean((El <= E) & (E <= Eh))))
