#!/usr/bin/env python
"""Gibbs iterations/s of the B200-native bayesNMF sampler (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of the R path

A "step" is one full Gibbs iteration (prior parameters -> P -> E -> latent counts ->
metrics; R/bayesNMF_sampler.R:273-285) on synthetic Poisson counts.  Default workload
"c3": Poisson-Gamma, K = 96, G = 100,000, N = 20 (BASELINE.json configs[2], the shape
the metric is quoted on), WGS-like counts (4,000 mutations per genome).  With N > 1 the
genomes are sharded over the ranks (one process per GPU, launched by torchrun) and the
K x N sufficient statistic is summed with NCCL each iteration: total work is fixed, so
scaling is "strong".  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (K, G, N, likelihood, prior, MH, learning_rank, mu_T)
    "c1": dict(K=96, G=100, N=5, likelihood="poisson", prior="gamma", MH=False, mu_T=4000.0),
    "c3": dict(K=96, G=100000, N=20, likelihood="poisson", prior="gamma", MH=False, mu_T=4000.0),
    "c3-exome": dict(K=96, G=100000, N=20, likelihood="poisson", prior="gamma", MH=False, mu_T=100.0),
}
# the other BASELINE.json configurations, timed by tools/time_configs.py (not bench lines)
METRIC = "Gibbs iterations/s (Poisson-Gamma, K=96, G=100000, N=20)"
L2_FLUSH_BYTES = 512 << 20


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel, workload, prec, world):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, from the committed
    `ncu --set full` capture of this workload (profiles/traffic.json names the capture); None
    when no capture of this exact configuration exists."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if world != 1 or not os.path.exists(p):
        return None
    e = json.load(open(p)).get(f"{kernel}:{workload}:{prec}")
    return None if e is None else e["bytes_per_launch"]


def ncu_limiters(kernel, workload, prec, world):
    """What the committed ncu capture says bounds `kernel` (issue slots, shared-memory pipe, DRAM)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if world != 1 or not os.path.exists(p):
        return None
    e = json.load(open(p)).get(f"{kernel}:{workload}:{prec}")
    return None if e is None else e.get("limiters")


def z_algorithmic_bytes(K, G, N, elem):
    """Bytes the fused latent-count kernel must move per launch: M read once (int32),
    E read once and SE written once, P read and SP written once."""
    return 4 * K * G + elem * N * G + 4 * N * G + elem * K * N + 8 * K * N


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, dev):
        self.rows, self.proc, self.dev = [], None, dev

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.dev)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            # gone before anything else is timed: a polling nvidia-smi holds the driver's lock for
            # milliseconds at a time, which shows up in every cudaMalloc / cudaMemcpy of this process
            self.proc.kill()
            self.proc.wait()
            self.th.join(timeout=2)

    def summary(self, t0, t1):
        rows = [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.15] or [r for _, r in self.rows]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in rows)
        reasons = []
        for i, nm in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]):
            if any(r[3 + i].lower().startswith("active") for r in rows):
                reasons.append(nm)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "reasons": reasons, "samples": len(rows)}


def synth(w, seed=0):
    from tests.util import synth_counts
    M, _, _ = synth_counts(w["K"], w["G"], w["N"], w["mu_T"], seed=seed)
    return M


# ------------------------------------------------------------------------------------------
def run_b200(args):
    # stdout carries exactly one JSON line: anything libraries print meanwhile (NCCL's version
    # banner, torchrun notices) goes to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    from bayesnmf_b200.shard import shard_bounds, sharded_handle

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w = dict(WORKLOADS[args.workload])
    K, G, N = w["K"], w["G"], w["N"]
    # the host buffer in the reference's own layout: R matrices are column-major doubles (REALSXP)
    M = np.asfortranarray(synth(w), dtype=np.float64)
    g_lo, g_hi = shard_bounds(G, rank, world)
    prec = args.precision
    elem = 8 if prec == "f64" else 4

    class _Solo:                      # the plumbing interface of torch.distributed for a 1-rank run
        @staticmethod
        def get_rank(): return 0
        @staticmethod
        def get_world_size(): return 1
        @staticmethod
        def is_initialized(): return False

    def make(ring_cap=0, share_comm=None):
        return sharded_handle(M, N, dist if world > 1 else _Solo, device=local, likelihood=w["likelihood"], prior=w["prior"],
                              MH=w["MH"], seed=1, precision=prec, ring_cap=ring_cap, share_comm=share_comm)

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    h = make()
    h.init_from_prior()
    h.set_l2_flush(L2_FLUSH_BYTES)
    h.step(args.warmup)
    sync()
    with ClockSampler(local) as cs:
        time.sleep(0.25)
        sync()
        t0 = time.time()
        out = h.step(args.steps)
        sync()
        t1 = time.time()
    tm = h.timing()
    clocks = cs.summary(t0, t1)
    iter_ms = tm["iter_ms"]
    if world > 1:
        t = torch.tensor([iter_ms, tm["zstat_ms"]], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        iter_ms, z_ms = float(t[0]), float(t[1])
    else:
        z_ms = tm["zstat_ms"]
    ms_per_step = iter_ms / args.steps
    value = 1e3 / ms_per_step

    # the same iterations back to back with a warm L2 (how a real chain runs)
    h.set_l2_flush(0)
    h.step(3)
    sync()
    h.step(args.steps)
    warm = h.timing()["iter_ms"]
    if world > 1:
        t = torch.tensor([warm], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        warm = float(t[0])
    last_row = out["metrics"][-1]
    # the latent-count kernel on its own (inside an iteration it shares the SMs with the side stream's
    # hyper-draws of the next iteration, which lengthens its launches and shortens the iteration)
    z_alone_ms = None
    if w["likelihood"] == "poisson" and not w["MH"]:
        zs = [h.sample_z(10_000 + i) for i in range(5)]
        z_alone_ms = float(np.median(zs[1:]))
        if world > 1:
            t = torch.tensor([z_alone_ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            z_alone_ms = float(t[0])

    # end to end through the public API from HOST buffers: construct (uploads the count
    # matrix), the prior draw, `steps` iterations with every sample_metrics row and every
    # P / A sample copied back, and the final E -- wall clock.  The NCCL communicator is the
    # process's existing one (a rendezvous is a once-per-process cost, not a per-run one).
    # One untimed pass first, as for the kernel timing: the sampler it closes leaves its device blocks in
    # the library's process-wide cache, as the previous rank's sampler does in a bayesNMF() call.
    hw = make(share_comm=h if world > 1 else None)
    hw.init_from_prior()
    hw.step(3, want_P=True, want_A=True)
    hw.get_state("E")
    hw.close()
    sync()
    e0 = time.time()
    h2 = make(share_comm=h if world > 1 else None)
    ea = time.time()
    h2.init_from_prior()
    o2 = h2.step(args.steps, want_P=True, want_A=True)
    eb = time.time()
    E_last = h2.get_state("E")
    ec = time.time()
    torch.cuda.synchronize()
    e1 = time.time()
    print(f"[e2e rank {rank}] construct+upload {ea - e0:.3f}s, prior draw + {args.steps} steps {eb - ea:.3f}s, final E {ec - eb:.3f}s, "
          f"device synchronize {e1 - ec:.3f}s", file=sys.stderr)
    h2.close()
    h.close()
    e2e_s = e1 - e0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t[0])
    h2d = (8 * K * (g_hi - g_lo) + 8 * 8) / args.steps
    d2h = (8 * 11 * (args.steps + 1) + 8 * K * N * args.steps + 8 * N * args.steps + 8 * N * (g_hi - g_lo)) / args.steps
    assert np.isfinite(o2["metrics"]).all() and np.isfinite(E_last).all()

    peak, peak_src = peaks()
    zb = z_algorithmic_bytes(K, g_hi - g_lo, N, elem)
    z_avg_ms = z_ms / args.steps
    achieved = zb / 1e9 / (z_avg_ms * 1e-3) if z_avg_ms > 0 else 0.0
    res = {
        "metric": METRIC, "value": value, "unit": "iterations/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": prec, "data": "synthetic",
        "config": {"workload": f"{args.workload}: {w['likelihood']}-{w['prior']} MH={w['MH']} K={K} G={G} N={N} "
                               f"mu_T={w['mu_T']:g} (sum M = {M.sum():.3g} latent picks/iteration)",
                   "sharding": f"G split over {world} rank(s), NCCL all-reduce of SP/rowSums(E)/metric partials",
                   "l2": f"flushed before every timed iteration ({L2_FLUSH_BYTES >> 20} MiB memset, outside the timed spans)",
                   "timing": "CUDA events per iteration on the sampler's stream, summed, max over ranks"},
        "value_l2_warm": 1e3 * args.steps / warm,
        "wall_ms_per_step_incl_flush": 1e3 * (t1 - t0) / args.steps,
        "gpu_launches": int(tm["launches"]),
        "clocks": clocks,
        "e2e": {"value": args.steps / e2e_s, "unit": "iterations/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "note": "handle creation + upload of M + prior draw + steps with all metric rows and P/A samples to host + final E; wall clock; one untimed pass of the same sequence (3 steps) before it"},
        "roofline": {"kernel": "k_zstat", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": ncu_traffic("k_zstat", args.workload, prec, world), "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": zb, "avg_launch_ms": z_avg_ms,
                     "launch_ms_kernel_alone": z_alone_ms,
                     "frac_kernel_alone": (zb / 1e9 / (z_alone_ms * 1e-3) / peak) if z_alone_ms else None,
                     "share_of_step": z_ms / iter_ms if iter_ms else None,
                     "limiters_ncu": ncu_limiters("k_zstat", args.workload, prec, world),
                     "latent_picks_per_s": float(M[:, g_lo:g_hi].sum()) / (z_avg_ms * 1e-3) if z_avg_ms > 0 else None},
        "last_metrics": {"RMSE": float(last_row[1]), "loglikelihood": float(last_row[3])},
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        res["cpu_baseline"] = cpu_baseline(args.workload, budget_s=args.cpu_budget)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    if rank == 0:
        print(json.dumps(res), flush=True)


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle (numpy restatement of the R path) on the box's host cores.
def _cpu_worker(a):
    name, g_lo, g_hi, G_sample, steps, mean_data = a
    import numpy as np  # noqa: F811
    from oracle.gibbs import OracleSampler
    w = WORKLOADS[name]
    M = synth(dict(w, G=G_sample))
    o = OracleSampler(M[:, g_lo:g_hi], w["N"], w["likelihood"], w["prior"], MH=w["MH"], seed=1,
                      g0=g_lo, G_total=G_sample, mean_data=mean_data)
    t0 = time.time()
    for _ in range(steps):
        o.step()
    return time.time() - t0


def cpu_oracle_rate(name, G_sample, steps, procs):
    """iterations/s of the full workload extrapolated from a G_sample-genome sample run on
    `procs` processes (each a contiguous genome shard; the per-genome cost is constant)."""
    import multiprocessing as mp
    w = WORKLOADS[name]
    M = synth(dict(w, G=G_sample))
    mean_data = float(M.mean())
    bounds = [(G_sample * i) // procs for i in range(procs + 1)]
    jobs = [(name, bounds[i], bounds[i + 1], G_sample, steps, mean_data) for i in range(procs)]
    ctx = mp.get_context("fork")
    t0 = time.time()
    with ctx.Pool(procs) as pool:
        times = pool.map(_cpu_worker, jobs)
    wall = time.time() - t0
    per_iter_sample = max(times) / steps
    per_iter_full = per_iter_sample * (w["G"] / G_sample)
    return 1.0 / per_iter_full, wall


def cpu_baseline(name, budget_s=20.0):
    procs = os.cpu_count() or 1
    w = WORKLOADS[name]
    G_sample = min(w["G"], 250 * procs)
    steps = 2
    rate, wall = cpu_oracle_rate(name, G_sample, steps, procs)
    return {"value": rate, "unit": "iterations/s", "cores": procs, "kind": "port",
            "sample": f"numpy oracle (fp64 restatement of the R sampler; R itself is not installed), {steps} iterations on "
                      f"{G_sample} of {w['G']} genomes split over {procs} processes, scaled by G ({wall:.1f} s of wall time)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    procs = os.cpu_count() or 1
    w = WORKLOADS[args.workload]
    G_sample = min(w["G"], 125 * procs)
    t_all = []
    cpu_oracle_rate(args.workload, G_sample, 1, procs) if args.warmup else None
    for _ in range(max(1, min(args.steps, 3))):
        rate, wall = cpu_oracle_rate(args.workload, G_sample, 1, procs)
        t_all.append(rate)
    value = float(np.median(t_all))
    sample = (f"numpy oracle (fp64 restatement of the R sampler; R is not installed on this image), "
              f"{len(t_all)} timed iteration(s) on {G_sample} of {w['G']} genomes over {procs} processes, scaled by G")
    res = {"impl": "reference", "metric": METRIC, "value": value, "unit": "iterations/s", "n_gpus": args.gpus,
           "steps": len(t_all), "warmup": 1 if args.warmup else 0, "ms_per_step": 1e3 / value, "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": f"{args.workload}: {w['likelihood']}-{w['prior']} MH={w['MH']} K={w['K']} G={w['G']} N={w['N']} mu_T={w['mu_T']:g}"},
           "cpu_baseline": {"value": value, "unit": "iterations/s", "cores": procs, "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(res), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="f64", choices=["f64", "f32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        args.warmup = max(args.warmup, 3)
        run_b200(args)


if __name__ == "__main__":
    main()
