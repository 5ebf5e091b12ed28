"""Torch-side helpers for the shard / gather / merge entry points (cpprob_sis_run_shard, _merge, _merge_padded), kept for
callers that bring their own transport and for the gloo tests of the host arithmetic.  The product's multi-GPU path no
longer goes through here: cpprob_sis_run_dist issues the all-gather itself (NCCL inside libcpprob_sis.so).

One process per GPU, particles sharded by whole chunks, and ONE small collective —
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


class _DeviceRows:
    """zero-copy view of an engine-owned device buffer"""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def gather_padded(p, n_total, world, cache):
    """The exchange of an inference with the fewest host-side steps: every rank contributes rows_per_rank = max over
    ranks of its row count, straight from the buffer cpprob_sis_run_shard returned (which always has room for one row
    more than it holds), in ONE all_gather_into_tensor; cpprob_sis_merge_padded compacts the segments on the device.
    p: the Partials of this rank.  cache: a dict the caller keeps between steps.  Returns (gathered tensor, rows_per_rank)."""
    key = (n_total, world, p.rows_per_chunk, p.n_cols)
    if cache.get("key") != key:
        m = max(shard_sizes(n_total, world, p.rows_per_chunk))
        dev = torch.device("cuda", torch.cuda.current_device())
        cache.update(key=key, m=m, out=torch.empty((world * m, p.n_cols), dtype=torch.float64, device=dev),
                     empty=torch.zeros((m, p.n_cols), dtype=torch.float64, device=dev))
    m = cache["m"]
    local = torch.as_tensor(_DeviceRows(p.device_ptr, (m, p.n_cols)), device="cuda") if p.n_chunks_local else cache["empty"]
    dist.all_gather_into_tensor(cache["out"], local)
    return cache["out"], m
