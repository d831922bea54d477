"""The N > 1 path on CPU: two `gloo` ranks run the sharded Gibbs iteration of the oracle with
the collectives at exactly the points where bnmf_step calls NCCL (SP, rowSums(E), metric
partials), through the same host plumbing the GPU path uses (bayesnmf_b200.shard).
Sharded == unsharded: integer statistics and P bit-identical, metrics to rounding."""
import os
import pickle
import socket
import tempfile

import numpy as np
import pytest

from bayesnmf_b200.shard import shard_bounds


def test_shard_bounds_partition():
    for G in (1, 7, 64, 100000):
        for world in (1, 2, 3, 8):
            b = [shard_bounds(G, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == G
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, out_dir, prior, learn):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torch
    from bayesnmf_b200.shard import broadcast_bytes, global_mean
    from oracle.gibbs import OracleSampler, get_temp_sched
    from tests.util import synth_counts
    M, _, _ = synth_counts(96, 37, 4, 900.0, seed=11)
    lo, hi = shard_bounds(M.shape[1], rank, world)

    def reduce_fn(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).clone()
        dist.all_reduce(t)
        return t.numpy()

    mean = global_mean(M[:, lo:hi], dist)
    mean_ls = global_mean(M[:, lo:hi], dist, local_sum=float(M[:, lo:hi].sum()))   # the sum bnmf_create hands back ('data_sum')
    uid = broadcast_bytes(bytes(range(128)) if rank == 0 else b"", 128, dist)     # how the NCCL id travels
    o = OracleSampler(M[:, lo:hi], 4, "poisson", prior, MH=False, seed=5, g0=lo, G_total=M.shape[1], mean_data=mean,
                      learning_rank=learn, temperature_schedule=get_temp_sched(40, 10) if learn else None, reduce_fn=reduce_fn)
    for _ in range(4):
        o.step()
    with open(os.path.join(out_dir, f"r{rank}.pkl"), "wb") as f:
        pickle.dump(dict(P=o.params["P"], E=o.params["E"], A=o.params["A"], SP=o.SP, SE=o.SE, metrics=o.metrics,
                         mean=mean, mean_ls=mean_ls, uid=uid, lo=lo, hi=hi), f)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("prior,learn", [("gamma", False), ("exponential", True)])
def test_two_rank_gloo_equals_single(prior, learn):
    import torch.multiprocessing as mp
    from oracle.gibbs import OracleSampler, get_temp_sched
    from tests.util import synth_counts
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(2, port, d, prior, learn), nprocs=2, join=True)
        parts = [pickle.load(open(os.path.join(d, f"r{r}.pkl"), "rb")) for r in range(2)]
    M, _, _ = synth_counts(96, 37, 4, 900.0, seed=11)
    ref = OracleSampler(M, 4, "poisson", prior, MH=False, seed=5, learning_rank=learn,
                        temperature_schedule=get_temp_sched(40, 10) if learn else None)
    for _ in range(4):
        ref.step()
    assert parts[0]["uid"] == parts[1]["uid"] == bytes(range(128))
    np.testing.assert_allclose(parts[0]["mean"], M.mean(), rtol=1e-14)
    assert parts[0]["mean_ls"] == parts[0]["mean"] == parts[1]["mean"] == parts[1]["mean_ls"]
    for p in parts:
        np.testing.assert_array_equal(p["SP"], ref.SP)                      # integer sums: exact
        np.testing.assert_array_equal(p["P"], ref.params["P"])              # replicated, bit-identical
        np.testing.assert_array_equal(p["A"], ref.params["A"])
        np.testing.assert_array_equal(p["E"], ref.params["E"][:, p["lo"]:p["hi"]])
        np.testing.assert_array_equal(p["SE"], ref.SE[:, p["lo"]:p["hi"]])
        for got, want in zip(p["metrics"], ref.metrics):
            for k in ("RMSE", "KL", "loglikelihood", "logposterior", "BIC", "rank"):
                np.testing.assert_allclose(got[k], want[k], rtol=1e-10, err_msg=k)
