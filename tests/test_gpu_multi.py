"""Genome-sharded run on 2 real GPUs (one process per GPU, NCCL sums inside bnmf_step) ==
the single-GPU run: SP and the replicated P bit-identical, E shard-wise identical, metric rows
to rounding.  Skipped on boxes with fewer than 2 GPUs (the driver's 1-GPU tier)."""
import os
import pickle
import socket
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out_dir, prior, learn, xchg="1"):
    import torch
    import torch.distributed as dist
    # BNMF_XCHG=0: the per-iteration sums through NCCL; default: the one-shot all-reduce over NVLink peer memory
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), BNMF_XCHG=xchg)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from bayesnmf_b200.shard import shard_bounds, sharded_handle
    from oracle.gibbs import get_temp_sched
    from tests.util import synth_counts
    M, _, _ = synth_counts(96, 1000, 6, 1200.0, seed=13)
    h = sharded_handle(M, 6, dist, device=rank, likelihood="poisson", prior=prior, MH=False, seed=8, learning_rank=learn)
    if learn:
        h.set_temperature_schedule(get_temp_sched(60, 12))
    rows = [h.init_from_prior()]
    out = h.step(12, want_P=True, want_A=True)
    lo, hi = shard_bounds(M.shape[1], rank, world)
    res = dict(P=h.get_state("P"), E=h.get_state("E"), SP=h.get_state("SP"), A=h.get_state("A"), metrics=out["metrics"],
               Ps=out["P"], lo=lo, hi=hi, row1=rows[0])
    with open(os.path.join(out_dir, f"r{rank}.pkl"), "wb") as f:
        pickle.dump(res, f)
    h.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("xchg", ["1", "0"])
@pytest.mark.parametrize("prior,learn", [("gamma", False), ("exponential", True)])
def test_two_gpu_shards_equal_single_gpu(built_lib, prior, learn, xchg):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from bayesnmf_b200 import Handle
    from oracle.gibbs import get_temp_sched
    from tests.util import synth_counts
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(2, port, d, prior, learn, xchg), nprocs=2, join=True)
        parts = [pickle.load(open(os.path.join(d, f"r{r}.pkl"), "rb")) for r in range(2)]
    M, _, _ = synth_counts(96, 1000, 6, 1200.0, seed=13)
    h = Handle(M, 6, likelihood="poisson", prior=prior, MH=False, seed=8, learning_rank=learn)
    if learn:
        h.set_temperature_schedule(get_temp_sched(60, 12))
    row1 = h.init_from_prior()
    out = h.step(12, want_P=True, want_A=True)
    P, E, SP, A = h.get_state("P"), h.get_state("E"), h.get_state("SP"), h.get_state("A")
    for p in parts:
        np.testing.assert_array_equal(p["SP"], SP)
        np.testing.assert_array_equal(p["P"], P)
        np.testing.assert_array_equal(p["Ps"], out["P"])
        np.testing.assert_array_equal(p["A"], A)
        np.testing.assert_array_equal(p["E"], E[:, p["lo"]:p["hi"]])
        np.testing.assert_allclose(p["metrics"], out["metrics"], rtol=1e-10)
        for k in ("RMSE", "KL", "loglikelihood", "logposterior"):
            np.testing.assert_allclose(p["row1"][k], row1[k], rtol=1e-10)
    np.testing.assert_array_equal(parts[0]["metrics"], parts[1]["metrics"])      # every rank returns the same rows


@pytest.mark.parametrize("prior,world", [("gamma", 2), ("exponential", 3)])
def test_shards_with_host_reduction_equal_single_handle(built_lib, prior, world):
    """The 1-GPU variant of the test above (the driver's 1-GPU tier skips that one): `world` shard
    handles of ONE process on one device, no communicator; after every iteration the host sums what
    the NCCL exchange of bnmf_step sums -- SP (exact integers) and rowSums(E) (exact fixed point) --
    and hands the totals back to every shard.  P, SP and the shards of E must equal the unsharded
    run bit for bit; the additive metric partials must add up."""
    from bayesnmf_b200 import Handle
    from bayesnmf_b200.hyperpriors import fill_hyperprior_params
    from bayesnmf_b200.shard import shard_bounds
    from tests.util import synth_counts
    K, G, N = 96, 1003, 6
    M, _, _ = synth_counts(K, G, N, 1200.0, seed=13)
    hyper = fill_hyperprior_params(None, prior, float(M.mean()), N)

    def make(lo, hi):
        h = Handle(M[:, lo:hi], N, likelihood="poisson", prior=prior, MH=False, seed=8, g0=lo, G_total=G)
        for k, v in hyper.items():
            h.set_hyper(k, v)
        return h

    whole = make(0, G)
    bounds = [shard_bounds(G, r, world) for r in range(world)]
    shards = [make(lo, hi) for lo, hi in bounds]

    def exchange():
        SP = sum(h.get_state("SP") for h in shards)
        rs = sum(h.get_state("rowsumE") for h in shards)
        for h in shards:
            h.set_state("SP", SP)
            h.set_state("rowsumE", rs)
        return SP

    row_w = whole.init_from_prior()
    rows = [h.init_from_prior() for h in shards]
    SP = exchange()
    np.testing.assert_array_equal(SP, whole.get_state("SP"))
    np.testing.assert_allclose(sum(r["loglikelihood"] for r in rows), row_w["loglikelihood"], rtol=1e-12)
    for it in range(8):
        mw = whole.step(1)["metrics"][0]
        ms = [h.step(1)["metrics"][0] for h in shards]
        SP = exchange()
        np.testing.assert_array_equal(SP, whole.get_state("SP"), err_msg=f"iteration {it}: SP")
        P = whole.get_state("P")
        E = whole.get_state("E")
        for h, (lo, hi) in zip(shards, bounds):
            np.testing.assert_array_equal(h.get_state("P"), P, err_msg=f"iteration {it}: P")
            np.testing.assert_array_equal(h.get_state("E"), E[:, lo:hi], err_msg=f"iteration {it}: E")
        np.testing.assert_allclose(sum(m[3] for m in ms), mw[3], rtol=1e-12)                  # log-likelihood adds up
        np.testing.assert_allclose(sum(m[2] for m in ms), mw[2], rtol=1e-9, atol=1e-6)        # so does the padded KL
    for h in shards + [whole]:
        h.close()
