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


def _worker(rank, world, port, out_dir, prior, learn):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
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


@pytest.mark.parametrize("prior,learn", [("gamma", False), ("exponential", True)])
def test_two_gpu_shards_equal_single_gpu(built_lib, prior, learn):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from bayesnmf_b200 import Handle
    from oracle.gibbs import get_temp_sched
    from tests.util import synth_counts
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(2, port, d, prior, learn), nprocs=2, join=True)
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
