#!/usr/bin/env python
"""Gibbs iterations/s of the B200-native bayesNMF sampler (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--workload c1|c2|c3|c3-exome|c4|c5]
    python bench.py --impl reference --gpus N --steps K --warmup W     # CPU arm (C++/OpenMP port; R is not installed)

A "step" is one full Gibbs iteration (prior parameters -> P -> E -> [R, A] -> latent counts | sigmasq ->
record -> metrics; R/bayesNMF_sampler.R:273-285) on synthetic data of the named shape.  Default workload
"c3": Poisson-Gamma, K = 96, G = 100,000, N = 20 (BASELINE.json configs[2], the shape the metric is quoted
on), WGS-like counts (4,000 mutations per genome).

Multi-GPU (one process per GPU under torchrun):
  c3, c3-exome   the genomes are SHARDED over the ranks, the K x N sufficient statistic is summed every
                 iteration (NCCL / NVLink): total work fixed => "scaling": "strong"
  c1 c2 c4 c5    REPLICAS: rank r runs the independent chain with seed r (BASELINE.json configs[1], [4]:
                 "one rank / chain per GPU"), no data-path collective => "scaling": "weak"
Prints ONE JSON line on rank 0.
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
    # BASELINE.json configs[0] .. [4] (SURVEY.md section 8, table of shapes) + the exome-like variant of C3
    "c1": dict(K=96, G=100, N=5, likelihood="poisson", prior="gamma", MH=False, learn=False, mu_T=4000.0, sharded=False,
               what="configs[0]: Poisson-Gamma, fixed rank 5, 96 x 100"),
    "c2": dict(K=96, G=500, N=10, likelihood="poisson", prior="truncnormal", MH=True, learn=True, mu_T=4000.0, sharded=False,
               what="configs[1]: Poisson-TruncNormal + MH, SBFI rank 0:10, 96 x 500"),
    "c3": dict(K=96, G=100000, N=20, likelihood="poisson", prior="gamma", MH=False, learn=False, mu_T=4000.0, sharded=True,
               what="configs[2]: Poisson-Gamma N = 20, 96 x 100,000, genomes sharded"),
    "c3-exome": dict(K=96, G=100000, N=20, likelihood="poisson", prior="gamma", MH=False, learn=False, mu_T=100.0, sharded=True,
                     what="configs[2] with exome-like counts (100 mutations per genome)"),
    "c4": dict(K=96, G=20000, N=15, likelihood="normal", prior="truncnormal", MH=False, learn=False, mu_T=4000.0, sharded=False,
               what="configs[3]: Normal-TruncNormal N = 15, 96 x 20,000"),
    "c5": dict(K=1536, G=50000, N=40, likelihood="poisson", prior="exponential", MH=True, learn=False, mu_T=4000.0, sharded=False,
               what="configs[4]: SBS1536, Poisson-Exponential + MH, N = 40, 1536 x 50,000, one chain per GPU"),
}
L2_FLUSH_BYTES = 512 << 20


def metric_name(w):
    model = f"{w['likelihood'].capitalize()}-{w['prior'].capitalize()}" + ("+MH" if w["MH"] else "") + (" SBFI" if w["learn"] else "")
    if w["likelihood"] == "poisson" and w["prior"] == "gamma" and w["G"] == 100000:
        return "Gibbs iterations/s (Poisson-Gamma, K=96, G=100000, N=20)"
    return f"Gibbs iterations/s ({model}, K={w['K']}, G={w['G']}, N={w['N']})"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return float(j["hbm_gbs"]), float(j.get("sm_max_mhz", 1965.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1965.0, "fallback (B200_PROFILING.md)"


def ncu_entry(kernel, workload, prec, world):
    """The committed `ncu --set full` capture of `kernel` at this workload (profiles/traffic.json): DRAM bytes and
    warp instructions per launch, what bounds the kernel.  None when no capture of this configuration exists."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if world != 1 or not os.path.exists(p):
        return None
    return json.load(open(p)).get(f"{kernel}:{workload}:{prec}")


def algorithmic_bytes(kernel, w, G, elem):
    """Bytes a kernel must move per launch: counts are int32, real data and the state are
    `elem` bytes.  G = genomes resident on this GPU."""
    K, N = w["K"], w["N"]
    dm = 4 if w["likelihood"] == "poisson" else elem
    KG, NG, KN = K * G, N * G, K * N
    table = {
        "k_zstat": 4 * KG + (elem + 4) * NG + (elem + 8) * KN,          # M, E in + SE out, P in + SP out
        "k_e_sweep": (dm + 2 * elem) * KG + 2 * elem * NG,               # M, Mhat in/out, E in/out
        "k_p_rows": (dm + 2 * elem) * KG + elem * NG,                    # M, Mhat in/out, E in
        "k_final": (dm + 2 * elem) * KG,
        "k_a_pass": (dm + 2 * elem) * KG,
        "k_a_sweep": N * (dm + 2 * elem) * KG,                           # the N passes of the rank learner in one launch
        "k_e_gram": dm * KG + 4 * elem * NG,                             # M once, E in/out + its two prior parameters
        "k_mhat_tc": elem * KG + elem * (KN + NG),                       # Mhat out, P and E in
        "k_mhat_full": elem * KG + elem * (KN + NG),
        "k_gram_part": dm * KG + elem * NG,
        "k_p_gram": elem * KN,
        "k_eside": (6 * elem + 4) * NG,
        "k_pside": (6 * elem + 8) * KN,
        "k_hyper": 4 * elem * NG,
    }
    return table.get(kernel)


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


def synth(w, seed=0, G=None):
    """Synthetic data of the workload's shape: M ~ Poisson(P_true E_true); the Normal
    workload adds N(0, (0.05 * column mean)^2) noise; rank learning plants half of the offered signatures."""
    from tests.util import synth_counts
    n_true = max(2, w["N"] // 2) if w["learn"] else w["N"]
    M, _, _ = synth_counts(w["K"], w["G"] if G is None else G, n_true, w["mu_T"], seed=seed)
    if w["likelihood"] == "normal":
        M = M + np.random.default_rng(1).normal(0, 0.05 * M.mean(axis=0, keepdims=True) + 1e-3, M.shape)
    return M


def temps_for(w):
    if not w["learn"]:
        return None
    from bayesnmf_b200.sampler import get_temp_sched
    return get_temp_sched(5000, 1000, np.random.default_rng(0))


# ------------------------------------------------------------------------------------------
def run_b200(args):
    # stdout carries exactly one JSON line: anything libraries print meanwhile (NCCL's version
    # banner, torchrun notices) goes to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    from bayesnmf_b200 import Handle
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
    sharded = w["sharded"]
    # the host buffer in the reference's own layout: R matrices are column-major doubles (REALSXP)
    M = np.asfortranarray(synth(w), dtype=np.float64)
    g_lo, g_hi = shard_bounds(G, rank, world) if sharded else (0, G)
    G_loc = g_hi - g_lo
    prec = args.precision
    elem = 8 if prec == "f64" else 4
    temps = temps_for(w)
    seed = 1 if sharded else 1 + rank            # replicas: chain r has its own seed

    class _Solo:                      # the plumbing interface of torch.distributed for a 1-rank run
        @staticmethod
        def get_rank(): return 0
        @staticmethod
        def get_world_size(): return 1
        @staticmethod
        def is_initialized(): return False

    def make(ring_cap=0, share_comm=None):
        if sharded:
            h = sharded_handle(M, N, dist if world > 1 else _Solo, device=local, likelihood=w["likelihood"], prior=w["prior"],
                               MH=w["MH"], seed=seed, precision=prec, ring_cap=ring_cap, share_comm=share_comm)
        else:
            h = Handle(M, N, likelihood=w["likelihood"], prior=w["prior"], MH=w["MH"], learning_rank=w["learn"], seed=seed,
                       precision=prec, device=local, ring_cap=ring_cap)
        if temps is not None:
            h.set_temperature_schedule(temps)
        return h

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def rmax(*vals):
        if world == 1:
            return [float(v) for v in vals]
        t = torch.tensor([float(v) for v in vals], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t]

    h = make()
    h.init_from_prior()
    h.set_l2_flush(L2_FLUSH_BYTES)
    h.step(args.warmup)
    sync()
    with ClockSampler(local) as cs:
        time.sleep(0.25)
        sync()
        t0 = time.time()
        os.environ["BNMF_TIMING"] = "iter"       # events at the iterations' boundaries only (none inside their chain)
        out = h.step(args.steps)
        os.environ.pop("BNMF_TIMING")
        sync()
        t1 = time.time()
    tm = h.timing()
    clocks = cs.summary(t0, t1)
    iter_ms, = rmax(tm["iter_ms"])
    ms_per_step = iter_ms / args.steps
    # the latent-count kernel's launches inside an iteration (events on either side of it): a pass of their own
    z_steps = min(args.steps, 20)
    h.step(z_steps)
    z_ms, = rmax(h.timing()["zstat_ms"])
    z_ms *= args.steps / z_steps
    n_chains = 1 if sharded else world
    value = n_chains * 1e3 / ms_per_step
    launches = int(tm["launches"])

    # the same iterations back to back with a warm L2 (how a real chain runs)
    h.set_l2_flush(0)
    h.step(3)
    sync()
    h.step(args.steps)
    warm, = rmax(h.timing()["iter_ms"])
    last_row = out["metrics"][-1]

    # Metropolis-Hastings models: the post-warm-up phase with the real accept step (R/sample_Pn.R:206-247)
    mh_phase = None
    if w["MH"]:
        h.set_l2_flush(L2_FLUSH_BYTES)
        h.step(3, converged=True)
        sync()
        o3 = h.step(args.steps, converged=True)
        mh_ms, = rmax(h.timing()["iter_ms"])
        mh_phase = {"value": n_chains * 1e3 * args.steps / mh_ms, "unit": "iterations/s", "ms_per_step": mh_ms / args.steps,
                    "P_mean_acceptance_rate": float(o3["metrics"][-1][9])}
        h.set_l2_flush(0)

    # per-kernel device time of an iteration (CUDA events after every launch, bnmf_profile_iteration): the
    # dominant kernel, its share of the step and its roofline
    h.set_l2_flush(L2_FLUSH_BYTES)
    prof = {}
    NPROF = 5
    for i in range(NPROF + 1):
        p = h.profile_iteration(converged=False)
        if i == 0:
            continue                  # (first one untimed)
        for k, (ms, c) in p.items():
            a = prof.setdefault(k, [0.0, 0])
            a[0] += ms; a[1] += c
    h.set_l2_flush(0)
    prof_total = sum(v[0] for v in prof.values())
    dom = max(prof, key=lambda k: prof[k][0])
    dom_ms = prof[dom][0] / prof[dom][1]
    dom_ms, = rmax(dom_ms)
    # the latent-count kernel on its own (inside an iteration it shares the SMs with the side stream's
    # hyper-draws of the next iteration, which lengthens its launches and shortens the iteration)
    z_alone_ms = None
    if w["likelihood"] == "poisson" and not w["MH"]:
        zs = [h.sample_z(10_000 + i) for i in range(5)]
        z_alone_ms, = rmax(float(np.median(zs[1:])))

    # end to end through the public API from HOST buffers: construct (uploads the data matrix), the prior
    # draw, `steps` iterations with every sample_metrics row and every P / A sample copied back, and the
    # final E -- wall clock.  The NCCL communicator is the process's existing one (a rendezvous is a
    # once-per-process cost, not a per-run one).  One untimed pass first, as for the kernel timing: the
    # sampler it closes leaves its device blocks in the library's process-wide cache, as the previous
    # rank's sampler does in a bayesNMF() call.
    e2e_steps = max(args.steps, args.e2e_steps)
    hw = make(share_comm=h if (world > 1 and sharded) else None)
    hw.init_from_prior()
    hw.step(3, want_P=True, want_A=True)
    hw.get_state("E")
    hw.close()
    sync()
    e0 = time.time()
    h2 = make(share_comm=h if (world > 1 and sharded) else None)
    ea = time.time()
    h2.init_from_prior()
    o2 = h2.step(e2e_steps, want_P=True, want_A=True)
    eb = time.time()
    E_last = h2.get_state("E")
    ec = time.time()
    torch.cuda.synchronize()
    e1 = time.time()
    print(f"[e2e rank {rank}] construct+upload {ea - e0:.4f}s, prior draw + {e2e_steps} steps {eb - ea:.4f}s, final E {ec - eb:.4f}s, "
          f"device synchronize {e1 - ec:.4f}s", file=sys.stderr)
    h2.close()
    h.close()
    e2e_s, = rmax(e1 - e0)
    h2d = (8 * K * G_loc + 8 * 8) / e2e_steps
    d2h = (8 * 11 * (e2e_steps + 1) + 8 * K * N * e2e_steps + 8 * N * e2e_steps + 8 * N * G_loc) / e2e_steps
    assert np.isfinite(o2["metrics"][:, :9]).all() and np.isfinite(E_last).all()

    peak, sm_max_mhz, peak_src = peaks()
    roof_kernel = "k_zstat" if (w["likelihood"] == "poisson" and not w["MH"] and sharded) else dom
    if roof_kernel == "k_zstat" and z_ms > 0:
        k_ms = z_ms / args.steps            # live launches inside the timed steps (under the side-stream overlap)
    else:
        k_ms = prof[roof_kernel][0] / prof[roof_kernel][1]
        k_ms, = rmax(k_ms)
    kb = algorithmic_bytes(roof_kernel, w, G_loc, elem)
    achieved = (kb / 1e9 / (k_ms * 1e-3)) if (kb and k_ms > 0) else None
    ncu = ncu_entry(roof_kernel, args.workload, prec, world)
    roofline = {"kernel": roof_kernel, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": ncu["bytes_per_launch"] if ncu else None,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": kb, "avg_launch_ms": k_ms,
                "share_of_step": prof[roof_kernel][0] / prof_total if roof_kernel in prof else None,
                "limiters_ncu": ncu.get("limiters") if ncu else None}
    if roof_kernel == "k_zstat":
        picks = float(M[:, g_lo:g_hi].sum())
        roofline.update({"launch_ms_kernel_alone": z_alone_ms,
                         "frac_kernel_alone": (kb / 1e9 / (z_alone_ms * 1e-3) / peak) if z_alone_ms else None,
                         "latent_picks_per_s": picks / (k_ms * 1e-3)})
        # the roof that governs this kernel: instruction issue (north star: "or FP32 issue rate where the step is
        # RNG-bound").  warp instructions per launch from the committed ncu capture / launch time, against
        # 148 SMs x 4 schedulers x 1 instruction per cycle at the clock seen during the run
        if ncu and ncu.get("warp_insts_per_launch") and z_alone_ms:
            f_hz = 1e6 * (clocks.get("sm_mhz") or sm_max_mhz)
            ipeak = 148 * 4 * f_hz
            iach = ncu["warp_insts_per_launch"] / (z_alone_ms * 1e-3)
            roofline["issue"] = {"bound": "issue", "achieved": iach / 1e9, "peak": ipeak / 1e9, "unit": "G warp-inst/s",
                                 "frac": iach / ipeak, "warp_insts_per_launch": ncu["warp_insts_per_launch"],
                                 "thread_insts_per_pick": 32.0 * ncu["warp_insts_per_launch"] / picks if picks else None,
                                 "source": "ncu smsp__inst_executed.sum (profiles/traffic.json) / CUDA-event time of the kernel alone"}
    if roof_kernel in ("k_p_rows", "k_e_sweep") and w["likelihood"] == "poisson" and k_ms:
        # the roof that governs the Poisson + MH sweeps: the fp64 pipe.  A cell (k, n, g) of a sweep costs one fp64
        # reciprocal (~11 instructions), the residual, two products into two sums and the rank-1 update of Mhat:
        # ~24 fp64 instructions, K N G cells per side
        f_hz = 1e6 * (clocks.get("sm_mhz") or sm_max_mhz)
        dpeak = 148 * 64 * f_hz
        dach = 24.0 * K * N * G_loc / (k_ms * 1e-3)
        roofline["fp64"] = {"bound": "fp64 pipe", "achieved": dach / 1e12, "peak": dpeak / 1e12, "unit": "T fp64-inst/s",
                            "frac": dach / dpeak, "fp64_insts_per_cell": 24,
                            "source": "algorithmic count (one reciprocal + residual + two sums + rank-1 update per cell and signature) / CUDA-event time of the kernel"}
    res = {
        "metric": metric_name(w), "value": value, "unit": "iterations/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if sharded else "weak",
        "vs_baseline": None, "dtype": prec, "data": "synthetic",
        "config": {"workload": f"{args.workload}: {w['what']}; {w['likelihood']}-{w['prior']} MH={w['MH']} rank learning={w['learn']} "
                               f"K={K} G={G} N={N} mu_T={w['mu_T']:g}" + (f" (sum M = {M.sum():.3g} latent picks/iteration)" if not w["MH"] and w["likelihood"] == "poisson" else ""),
                   "sharding": (f"G split over {world} rank(s), all-reduce of SP/rowSums(E)/metric partials every iteration" if sharded
                                else f"{world} independent chain(s), one per GPU (seeds 1..{world}); value = chains x iterations/s of the slowest"),
                   "l2": f"flushed before every timed iteration ({L2_FLUSH_BYTES >> 20} MiB memset, outside the timed spans)",
                   "timing": "CUDA events per iteration on the sampler's stream, summed, max over ranks",
                   "phase": "warm-up iterations (every proposal accepted, R/sample_Pn.R:201-204); the real MH accept step is `mh_phase`" if w["MH"] else "every iteration is the same"},
        "value_l2_warm": n_chains * 1e3 * args.steps / warm,
        "wall_ms_per_step_incl_flush": 1e3 * (t1 - t0) / args.steps,
        "gpu_launches": launches,
        "clocks": clocks,
        "e2e": {"value": n_chains * e2e_steps / e2e_s, "unit": "iterations/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps,
                "note": "handle creation + upload of M + prior draw + steps with all metric rows and P/A samples to host + final E; wall clock; one untimed pass of the same sequence (3 steps) before it"},
        "roofline": roofline,
        "kernels_ms_per_step": {k: round(v[0] / NPROF, 5) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])},
        "last_metrics": {"RMSE": float(last_row[1]), "loglikelihood": float(last_row[3])},
    }
    if mh_phase:
        res["mh_phase"] = mh_phase
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
# CPU arm.  R is not installed in this image, so "the reference on the host cores" is a port:
#   Poisson latent-count models (c1, c3): oracle/cpu_port.cpp, C++ / OpenMP on every core, the FULL workload;
#   sweep models (c2, c4, c5): the numpy oracle (oracle/gibbs.py), the full workload where an iteration takes
#   seconds, else a stated genome sample scaled by G.
def cpu_port_run(name, steps, warmup):
    from oracle.cpu_port import CpuPort, lib
    w = WORKLOADS[name]
    M = synth(w)
    c = CpuPort(M, w["N"], w["prior"], seed=1)
    c.init_from_prior()
    if warmup:
        c.step(warmup)
    times = []
    for _ in range(steps):
        t0 = time.time()
        c.step(1)
        times.append(time.time() - t0)
    c.close()
    cores = int(lib().cp_threads())
    return 1.0 / float(np.mean(times)), cores, (f"C++/OpenMP port of the R sampler (oracle/cpu_port.cpp, fp64; R itself is not installed), the full workload "
                                                f"(all {w['G']} genomes), {steps} timed + {warmup} warm-up iteration(s) on {cores} threads, {np.mean(times):.2f} s per iteration")


def _numpy_worker(a):
    name, G_sample, steps, warmup = a
    from oracle.gibbs import OracleSampler
    w = WORKLOADS[name]
    M = synth(w, G=G_sample)
    o = OracleSampler(M, w["N"], w["likelihood"], w["prior"], MH=w["MH"], seed=1, learning_rank=w["learn"],
                      temperature_schedule=temps_for(w))
    for _ in range(warmup):
        o.step()
    t0 = time.time()
    for _ in range(steps):
        o.step()
    return (time.time() - t0) / steps


def numpy_oracle_run(name, steps, warmup):
    w = WORKLOADS[name]
    G_sample = w["G"] if w["K"] * w["G"] * w["N"] <= 40_000_000 else max(256, 40_000_000 // (w["K"] * w["N"]))
    per_iter = _numpy_worker((name, G_sample, steps, warmup)) * (w["G"] / G_sample)
    cores = os.cpu_count() or 1
    what = "the full workload" if G_sample == w["G"] else f"{G_sample} of {w['G']} genomes, scaled by G"
    return 1.0 / per_iter, cores, (f"numpy oracle (oracle/gibbs.py: fp64 restatement of the R sampler; R itself is not installed), {what}, "
                                   f"{steps} timed + {warmup} warm-up iteration(s), BLAS threads of {cores} cores")


def cpu_arm(name, steps, warmup):
    w = WORKLOADS[name]
    if w["likelihood"] == "poisson" and not w["MH"] and not w["learn"]:
        rate, cores, sample = cpu_port_run(name, steps, warmup)
        return {"value": rate, "unit": "iterations/s", "cores": cores, "kind": "port-c++", "sample": sample}
    rate, cores, sample = numpy_oracle_run(name, steps, warmup)
    return {"value": rate, "unit": "iterations/s", "cores": cores, "kind": "port", "sample": sample}


def cpu_baseline(name, budget_s=20.0):
    """A bounded sample (about `budget_s` seconds) of the CPU arm."""
    w = WORKLOADS[name]
    heavy = w["K"] * w["G"] * w["N"] > 50_000_000
    return cpu_arm(name, steps=3 if heavy else 10, warmup=1)


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    b = cpu_arm(args.workload, steps=max(1, args.steps), warmup=max(0, args.warmup))
    value = b["value"]
    res = {"impl": "reference", "metric": metric_name(w), "value": value, "unit": "iterations/s", "n_gpus": args.gpus,
           "steps": max(1, args.steps), "warmup": max(0, args.warmup), "ms_per_step": 1e3 / value, "higher_is_better": True,
           "scaling": "strong" if w["sharded"] else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": f"{args.workload}: {w['what']}; {w['likelihood']}-{w['prior']} MH={w['MH']} rank learning={w['learn']} "
                                  f"K={w['K']} G={w['G']} N={w['N']} mu_T={w['mu_T']:g}"},
           "cpu_baseline": b,
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
    ap.add_argument("--e2e-steps", type=int, default=100,
                    help="iterations of the end-to-end pass (at least --steps): construction + upload are paid once per run")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.steps == 100 and args.warmup == 5 and WORKLOADS[args.workload]["G"] * WORKLOADS[args.workload]["K"] > 5_000_000:
            args.steps, args.warmup = 10, 1          # the no-flag default stays within a few minutes
        run_reference(args)
    else:
        args.warmup = max(args.warmup, 3)
        run_b200(args)


if __name__ == "__main__":
    main()
