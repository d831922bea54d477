"""Replica placement ("replicas only"): independent chains, or the independent
fixed-rank samplers of a `rank_method = "BIC"` range (R/bayesNMF.R:66-127), one per GPU.

The reference runs them one after the other (`lapply(rank, ...)`, R/bayesNMF.R:67: serial).  Nothing
is exchanged between replicas, so there is no collective: every replica is a `bnmf_handle` on its
own device with its own stream, driven by a host thread of ONE process (the C ABI releases the
interpreter lock while a call runs; `cfg.device` of include/bnmf.h places the handle).  Replica i
runs on device `devices[i % len(devices)]`; a replica's chain is a pure function of (data, model,
seed) -- Philox is keyed by the seed -- so it is bit-identical to the same chain run alone on any
GPU (tests/test_gpu_replicas.py).  `bench.py --workload c5 --gpus 8` uses the other launcher the
contract prescribes (one process per GPU under torchrun, no data-path collective)."""
import threading

import numpy as np


def visible_devices():
    """CUDA device ordinals of this process (torch is plumbing: it only counts them)."""
    import torch
    return list(range(torch.cuda.device_count()))


def place(n_replicas, devices=None):
    """Device of every replica: round-robin over `devices` (default: every visible GPU)."""
    devices = visible_devices() if devices is None else list(devices)
    if not devices:
        raise RuntimeError("bayesnmf_b200.replicas: no CUDA device visible (there is no CPU path)")
    return [devices[i % len(devices)] for i in range(n_replicas)]


def run_replicas(jobs, devices=None):
    """Run `jobs` -- callables taking `device=` -- concurrently, at most one per device at a time
    (a device's replicas queue behind one another, the devices run side by side).  Returns the
    results in job order; the first exception of any replica is re-raised."""
    where = place(len(jobs), devices)
    out, err = [None] * len(jobs), []
    by_dev = {}
    for i, dev in enumerate(where):
        by_dev.setdefault(dev, []).append(i)

    def worker(dev, idx):
        for i in idx:
            if err:
                return
            try:
                out[i] = jobs[i](device=dev)
            except BaseException as e:      # noqa: BLE001 -- handed to the caller's thread
                err.append(e)
                return

    threads = [threading.Thread(target=worker, args=(dev, idx), daemon=True) for dev, idx in by_dev.items()]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if err:
        raise err[0]
    return out


def run_chains(data, rank, seeds, devices=None, **kw):
    """Independent chains of one model (BASELINE.json configs[4]: "8 independent chains on 8 x B200"):
    chain i = bayesNMF(data, rank, seed = seeds[i], ...) on its own GPU.  Returns the samplers."""
    from .sampler import bayesNMF_sampler

    def job(seed):
        return lambda device: bayesNMF_sampler(data, rank, seed=int(seed), device=device, **kw).run_gibbs_sampler()

    return run_replicas([job(s) for s in seeds], devices)


def step_chains(data, N, seeds, n_iters, devices=None, converged=False, **kw):
    """The bare hot path for `len(seeds)` chains: handle, prior draw, `n_iters` iterations each;
    returns [(metrics rows, P)] per chain.  (What the replica test and the bench time.)"""
    from . import Handle

    def job(seed):
        def run(device):
            h = Handle(np.asarray(data, dtype=np.float64), N, seed=int(seed), device=device, **kw)
            try:
                h.init_from_prior()
                met = h.step(n_iters, converged=converged)["metrics"]
                return met, h.get_state("P")
            finally:
                h.close()
        return run

    return run_replicas([job(s) for s in seeds], devices)
