"""Genome sharding for one-process-per-GPU runs.

The count matrix is split into contiguous column blocks; every rank owns one `Handle` for its
block (cfg.G = local columns, cfg.g0 = first global column, cfg.G_total).  `torch.distributed`
is plumbing only: it carries the NCCL unique id and the global data mean to every rank; the
per-iteration sums (SP, rowSums(E), metric partials) are NCCL calls inside bnmf_step.
Works with any backend for the plumbing ("nccl" on the GPU box, "gloo" in the CPU tests).
"""
import numpy as np


def shard_bounds(G, rank, world):
    """Columns [lo, hi) of rank `rank`: contiguous, disjoint, covering 0..G, sizes differ by <= 1."""
    return (G * rank) // world, (G * (rank + 1)) // world


def _dev(dist):
    import torch
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def global_mean(M_local, dist=None, local_sum=None):
    """mean(data) over all shards (the hyperprior defaults of R/setup.R:123-181 depend on it).  `local_sum`: the
    shard's sum when the caller has it already (bnmf_create adds the data up while it converts them)."""
    s = float(np.sum(M_local, dtype=np.float64)) if local_sum is None else float(local_sum)
    n = float(np.size(M_local))
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return s / n
    import torch
    t = torch.tensor([s, n], dtype=torch.float64, device=_dev(dist))
    dist.all_reduce(t)
    return float(t[0]) / float(t[1])


def broadcast_bytes(payload, nbytes, dist, src=0):
    """Rank `src` supplies `payload` (bytes of length nbytes); every rank returns it."""
    import torch
    t = torch.zeros(nbytes, dtype=torch.uint8, device=_dev(dist))
    if dist.get_rank() == src:
        t.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    dist.broadcast(t, src)
    return bytes(t.cpu().numpy().tobytes())


def sharded_handle(M, N, dist, device=0, hyperprior_params=None, share_comm=None, **kw):
    """Handle for this rank's block of the full matrix M (K x G_total), joined to the NCCL
    communicator of all ranks (a new one, or the one `share_comm` -- another handle of this
    process -- already holds), with the hyperprior defaults of the *whole* data set."""
    from . import Handle, comm_unique_id
    from .hyperpriors import fill_hyperprior_params
    rank, world = dist.get_rank(), dist.get_world_size()
    G = M.shape[1]
    lo, hi = shard_bounds(G, rank, world)
    h = Handle(M[:, lo:hi], N, device=device, g0=lo, G_total=G, **kw)
    if world > 1 or hyperprior_params:
        # (a single unsharded handle without user values keeps the defaults bnmf_create installed from
        #  the mean of its own -- that is, of all -- columns: R/setup.R:123-181)
        mean = global_mean(M[:, lo:hi], dist, local_sum=h.get_state("data_sum")[0])
        for name, value in fill_hyperprior_params(hyperprior_params, kw.get("prior", "gamma"), mean, N).items():
            if np.ndim(value) == 2 and name.endswith("_e"):
                value = np.asarray(value)[:, lo:hi]
            h.set_hyper(name, value)
    if world > 1 and share_comm is not None:
        h.comm_share(share_comm)
    elif world > 1:
        uid = broadcast_bytes(comm_unique_id() if rank == 0 else b"", 128, dist)
        h.comm_init(uid, rank, world)
    return h
