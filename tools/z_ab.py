"""A/B timing of k_zstat alone over experiment builds (exp/lib_*.so; '-' = the product library):
   python tools/z_ab.py <mu> <G,G,...> <lib> [<lib> ...]   (dev tool; the counts are cached under /tmp)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CHILD = r'''
import sys, numpy as np
sys.path.insert(0, %r)
from bayesnmf_b200 import Handle
from bayesnmf_b200.hyperpriors import fill_hyperprior_params
M = np.load(sys.argv[1])
h = Handle(M, 20, likelihood="poisson", prior="gamma", MH=False, seed=1)
for k, v in fill_hyperprior_params(None, "gamma", float(M.mean()), 20).items():
    h.set_hyper(k, v)
h.init_from_prior(); h.step(2)
z = [h.sample_z(50 + i) for i in range(6)]
print(round(min(z), 4), round(float(np.median(z)), 4), float(h.get_state("SP").sum()))
''' % ROOT
if __name__ == "__main__":
    mu = float(sys.argv[1]); Gs = [int(g) for g in sys.argv[2].split(",")]; libs = sys.argv[3:]
    from tests.util import synth_counts
    for G in Gs:
        f = f"/tmp/zab_{int(mu)}_{G}.npy"
        if not os.path.exists(f):
            np_M = synth_counts(96, G, 20, mu, seed=0)[0]
            import numpy as np
            np.save(f, np_M)
        for lib in libs:
            env = dict(os.environ)
            lib, _, kv = lib.partition("@")                      # lib@ENV=value,ENV=value: tuning knobs of the library
            for e in filter(None, kv.split(",")):
                env[e.split("=")[0]] = e.split("=")[1]
            if lib != "-": env["BNMF_LIB"] = os.path.join(ROOT, lib)
            else: env.pop("BNMF_LIB", None)
            r = subprocess.run([sys.executable, "-c", CHILD, f], env=env, capture_output=True, text=True, timeout=300)
            print(f"mu={mu:g} G={G} lib={lib} {kv}: min/median ms, sum(SP) = {r.stdout.strip() or r.stderr.strip()[-300:]}", flush=True)
