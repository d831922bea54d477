"""Replica placement (bayesnmf_b200/replicas.py): chains / BIC ranks one per GPU, driven concurrently by
the threads of one process.  A replica equals the same chain run alone (on a 1-GPU box the replicas
share device 0 and still run from concurrent host threads; on a multi-GPU box they spread)."""
import numpy as np
import pytest

from tests.util import synth_counts

pytestmark = pytest.mark.gpu


def test_chain_on_its_gpu_equals_the_chain_alone(built_lib):
    import torch
    from bayesnmf_b200 import Handle
    from bayesnmf_b200.replicas import place, step_chains
    M, _, _ = synth_counts(96, 200, 5, 1500.0, seed=6)
    seeds = [0, 1, 2, 3]
    ndev = torch.cuda.device_count()
    where = place(len(seeds))
    assert sorted(set(where)) == list(range(min(ndev, len(seeds))))
    kw = dict(likelihood="poisson", prior="exponential", MH=True)
    got = step_chains(M, 5, seeds, 6, **kw)
    for s, (met, P) in zip(seeds, got):
        h = Handle(M, 5, seed=s, device=0, **kw)
        h.init_from_prior()
        ref = h.step(6)["metrics"]
        np.testing.assert_array_equal(met, ref, err_msg=f"chain {s}")
        np.testing.assert_array_equal(P, h.get_state("P"), err_msg=f"chain {s}")
        h.close()
    assert not np.array_equal(got[0][1], got[1][1])          # different seeds, different chains


def test_bic_ranks_one_per_gpu_equal_the_serial_fan_out(built_lib):
    """R/bayesNMF.R:66-127: one fixed-rank sampler per rank; placed over the devices it returns the
    BIC table and best rank of the serial loop."""
    import torch
    from bayesnmf_b200.sampler import bayesNMF, new_convergence_control
    M, _, _ = synth_counts(96, 60, 3, 1500.0, seed=9)
    cc = new_convergence_control(MAP_over=20, MAP_every=10, miniters=20, maxiters=60)
    kw = dict(likelihood="poisson", prior="gamma", rank_method="BIC", convergence_control=cc, seed=4)
    serial = bayesNMF(M, [2, 3, 4], **kw)
    devs = list(range(torch.cuda.device_count())) * 2         # >= 2 entries: the threaded path even on one GPU
    spread = bayesNMF(M, [2, 3, 4], devices=devs, **kw)
    assert [r["rank"] for r in spread["results"]] == [r["rank"] for r in serial["results"]]
    np.testing.assert_array_equal([r["BIC"] for r in spread["results"]], [r["BIC"] for r in serial["results"]])
    assert spread["best_rank"] == serial["best_rank"]
    serial["sampler"].close(); spread["sampler"].close()
