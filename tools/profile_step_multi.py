#!/usr/bin/env python3
"""Host-side breakdown of one multi-GPU inference step (torchrun): run_shard / gather / merge wall times, rank 0.
usage: python -m torch.distributed.run --nproc-per-node N tools/profile_step_multi.py"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import bench
from cpprob_b200 import Engine
from cpprob_b200.dist import gather_partials, shard_sizes

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
e = Engine(device=local, seed=bench.SEED)
n = 1_000_000_000 * world
acc = {"shard": 0.0, "gather": 0.0, "sync": 0.0, "merge": 0.0}
scratch = None
for it in range(25):
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    p = e.run_shard(bench.MODEL, bench.OBS, n, rank, world)
    t1 = time.perf_counter()
    loc = torch.as_tensor(bench._DeviceArray(p.device_ptr, (p.n_chunks_local, p.n_cols)), device="cuda")
    if scratch is None:
        m = max(shard_sizes(n, world, p.rows_per_chunk))
        scratch = (torch.zeros((m, p.n_cols), dtype=torch.float64, device="cuda"), torch.empty((world, m, p.n_cols), dtype=torch.float64, device="cuda"))
    g = gather_partials(loc, n, world, scratch, p.rows_per_chunk)
    t2 = time.perf_counter()
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    st, _ = e.merge(bench.MODEL, bench.OBS, g.data_ptr(), g.shape[0], p.n_cols, p.m_ref, n)
    t4 = time.perf_counter()
    if it >= 5:
        acc["shard"] += t1 - t0; acc["gather"] += t2 - t1; acc["sync"] += t3 - t2; acc["merge"] += t4 - t3
if rank == 0:
    print({k: round(v / 20 * 1e3, 4) for k, v in acc.items()}, "ms per step; kernel", p.device_ms, "rows", g.shape)
dist.destroy_process_group()
