"""The reference's only published timing (vignettes/advanced.pdf p.9-10): bayesNMF(data$M, 1:10)
with every default (Poisson / truncated normal + MH, SBFI, maxiters 5000, MAP_over 1000,
post_warmup 2000) on the bundled 96 x 64 example took 15.99 min for 4,100 iterations = 4.27 it/s
on unstated hardware.  Same call through the host mirror, on the GPU."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bayesnmf_b200 import bayesNMF
from tests.util import example_data

M, Ptrue = example_data()
out = []
for seed in (1, 2, 3):
    t0 = time.time()
    s = bayesNMF(M, np.arange(1, 11), seed=seed, save_all_samples=False)
    dt = time.time() - t0
    c = (s.MAP["P"].T @ Ptrue) / np.outer(np.linalg.norm(s.MAP["P"], axis=0), np.linalg.norm(Ptrue, axis=0))
    out.append(dict(seed=seed, iterations=s.state["iter"], converged_iter=s.state.get("converged_iter"), why=s.state.get("why"),
                    seconds=round(dt, 3), it_per_s=round(s.state["iter"] / dt, 1), rank=int(s.MAP["P"].shape[1]),
                    cosine_to_planted=[round(float(x), 4) for x in c.max(axis=0)]))
    s.close()
    print(json.dumps(out[-1]), flush=True)
# the same run with the driver loop behind the ABI (bnmf_run): construction, one call, MAP + credible intervals
from bayesnmf_b200 import bayesNMF_sampler
dev = []
for seed in (1, 2, 3):
    t0 = time.time()
    s = bayesNMF_sampler(M, np.arange(1, 11), seed=seed, save_all_samples=False)
    r = s._h.run(s.specs["convergence_control"], post_warmup=s.specs["post_warmup"])
    n_s = min(s.specs["convergence_control"]["MAP_over"], s._h.ring_count())
    P_map, E_map, A_map, nm = s._h.get_map(n_s)
    ci = s._h.get_credible_intervals(n_s)
    dt = time.time() - t0
    keep = np.nonzero(A_map == 1)[0]
    c = (P_map[:, keep].T @ Ptrue) / np.outer(np.linalg.norm(P_map[:, keep], axis=0), np.linalg.norm(Ptrue, axis=0))
    dev.append(dict(seed=seed, iterations=r["iter"], converged_iter=r["converged_iter"], why=r["why"], seconds=round(dt, 3),
                    it_per_s=round(r["iter"] / dt, 1), rank=int(len(keep)), cosine_to_planted=[round(float(x), 4) for x in c.max(axis=0)]))
    s.close()
    print(json.dumps(dict(device_loop=dev[-1])), flush=True)
print(json.dumps(dict(reference_published=dict(iterations=4100, minutes=15.99, it_per_s=4.27,
                                               cosine=[0.9993, 0.9642, 0.9993, 0.9996], rank=4), runs=out, runs_device_loop=dev)))
