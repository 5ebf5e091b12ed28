"""Multi-GPU plumbing: one process per GPU, particles sharded by whole chunks, and ONE small collective —
an all-gather of the per-(super-)chunk partial sums in rank order (NCCL over NVLink on the GPU box, gloo in the CPU
tests).  Merging the gathered rows in chunk order (cpprob_sis_merge) gives bit-identical results on every
rank and for every world size; there is no data-path collective because particles are i.i.d.
(/root/reference include/cpprob/cpprob.hpp:194-201 has no inter-particle dependence)."""
import functools

import torch
import torch.distributed as dist

from . import capi


@functools.lru_cache(maxsize=64)
def shard_sizes(n_total, world, rows_per_chunk=1):
    """partial rows of every rank (host arithmetic of cpprob_sis_plan_rows): chunk rows for runs of up to 4096
    chunks, super-chunk rows (2^k chunks each, reduced on the owning rank) beyond that — never more than 4096 in all."""
    return tuple(capi.plan_rows(n_total, r, world, rows_per_chunk)[1] for r in range(world))


def gather_partials(local, n_total, world, scratch=None, rows_per_chunk=1):
    """local: [n_chunks_local, n_cols] float64 tensor of this rank.  Returns [n_chunks_total, n_cols] with
    every rank's rows in chunk order.  Shards differ by at most one chunk, so rows are padded to the
    largest shard for a single all_gather_into_tensor."""
    sizes = shard_sizes(n_total, world, rows_per_chunk)
    if world == 1:
        return local
    max_local = max(sizes)
    n_cols = local.shape[1]
    if scratch is None or scratch[0].shape != (max_local, n_cols) or scratch[0].device != local.device:
        scratch = (torch.zeros((max_local, n_cols), dtype=torch.float64, device=local.device),
                   torch.empty((world, max_local, n_cols), dtype=torch.float64, device=local.device))
    padded, allbuf = scratch
    padded[:local.shape[0]].copy_(local)
    dist.all_gather_into_tensor(allbuf.view(world * max_local, n_cols), padded)
    return torch.cat([allbuf[r, :sizes[r]] for r in range(world)]).contiguous()
