"""Per-iteration device time of an 8-GPU shard's iteration on ONE GPU (12,500 genomes of the C3 matrix, no exchange,
graphs off so that the launch path is the sharded one), alternating launch variants in one process (dev tool):
   BNMF_GRAPH=0 python tools/shard_ab.py [G] [iters]"""
import os, sys
os.environ.setdefault("BNMF_GRAPH", "0")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bayesnmf_b200 import Handle
from bayesnmf_b200.hyperpriors import fill_hyperprior_params
from tests.util import synth_counts

G = int(sys.argv[1]) if len(sys.argv) > 1 else 12500
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 300
M, _, _ = synth_counts(96, G, 20, 4000.0, seed=0)
h = Handle(M, 20, likelihood="poisson", prior="gamma", MH=False, seed=1)
for k, v in fill_hyperprior_params(None, "gamma", float(M.mean()), 20).items():
    h.set_hyper(k, v)
h.init_from_prior()
h.step(50)
variants = {"default": {}, "no timing events": {"BNMF_TIMING": "0"}, "z events too": {"BNMF_TIMING": "z"}, "two launches": {"BNMF_SIDES": "0"}, "k_sides forced": {"BNMF_SIDES": "1"}}
for a in sys.argv[3:]:                                  # extra variants: name:ENV=v,ENV=v
    nm, _, kv = a.partition(":")
    variants[nm] = dict(e.split("=") for e in kv.split(","))
KNOBS = sorted({k for v in variants.values() for k in v})
res = {k: [] for k in variants}
zin = {}
for rep in range(6):
    for name, env in variants.items():
        for k in KNOBS:
            os.environ.pop(k, None)
        os.environ.update(env)
        h.step(iters)
        res[name].append(h.timing()["total_ms"] / iters * 1e3)
        zin[name] = h.timing()["zstat_ms"] / iters * 1e3
for name, v in res.items():
    print(f"G={G} {name:15s}: us/iteration min {min(v):.2f} median {np.median(v):.2f}  all {np.round(v, 1)}  (k_zstat in-step {zin[name]:.1f})")
print("z alone us:", round(1e3 * min(h.sample_z(10 + i) for i in range(4)), 1))
