"""N>1 host-side logic on CPU: world_size-2 (3, 5) gloo groups shard the chunks, all-gather their partial rows and must end
up with the same, chunk-ordered matrix on every rank — both through the torch-side helper (cpprob_b200/dist.py) and with the
padded single all-gather + in-place row lookup that the library's own NCCL path uses (cpprob_sis_run_dist,
k_merge_columns_gathered).  The NCCL path itself runs in tests/test_dist_gpu.py (-m gpu)."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %r)
    import torch, torch.distributed as dist
    from cpprob_b200 import capi
    from cpprob_b200.dist import gather_partials, shard_sizes
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n_total = 11 * capi.CHUNK + 5          # 12 chunks: uneven over 5 ranks, even over 2/3/4
    n_cols = 10
    cf, ncl, nct, fp, nl = capi.plan_shard(n_total, rank, world)
    local = torch.empty((ncl, n_cols), dtype=torch.float64)
    for i in range(ncl):
        local[i] = torch.arange(n_cols, dtype=torch.float64) + 100.0 * (cf + i)    # row content = f(global chunk)
    g = gather_partials(local, n_total, world)
    expect = torch.stack([torch.arange(n_cols, dtype=torch.float64) + 100.0 * c for c in range(nct)])
    assert g.shape == (nct, n_cols) and torch.equal(g, expect), (rank, g)
    assert sum(shard_sizes(n_total, world)) == nct
    # every rank holds the identical matrix: a checksum all-reduce(max) == all-reduce(min)
    s = g.sum().reshape(1).clone(); lo = s.clone(); hi = s.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert lo.item() == hi.item()
    # The exchange the library performs (cpprob_sis_run_dist): every rank contributes rows_per_rank = the largest shard's
    # row count, padding included, in ONE all-gather; the merge kernel then finds logical row i at segment r = its owner,
    # position i - first[r], from cpprob_sis_plan_rows alone (k_merge_columns_gathered).  Same arithmetic, on the CPU:
    for n_big, per_chunk in ((n_total, 1), (n_total, 8), (5000 * capi.CHUNK + 777, 1), (5000 * capi.CHUNK + 777, 8)):
        plan = [capi.plan_rows(n_big, r, world, per_chunk) for r in range(world)]           # (first, n_local, n_total)
        first = [p[0] for p in plan] + [plan[-1][2]]
        assert first[0] == 0 and all(first[r] + plan[r][1] == first[r + 1] for r in range(world))
        rows_per_rank = max(p[1] for p in plan)
        mine = torch.full((rows_per_rank, 3), float("nan"), dtype=torch.float64)         # padding is never read
        for i in range(plan[rank][1]):
            mine[i] = float(first[rank] + i)                                                # row content = its logical index
        gathered = torch.empty((world * rows_per_rank, 3), dtype=torch.float64)
        dist.all_gather_into_tensor(gathered, mine)
        r_own = 0
        for i in range(first[-1]):
            while i >= first[r_own + 1]:
                r_own += 1
            phys = r_own * rows_per_rank + (i - first[r_own])
            assert gathered[phys, 0].item() == float(i), (rank, i, phys)
    dist.destroy_process_group()
    print("rank", rank, "ok")
""") % ROOT


@pytest.mark.parametrize("world", [2, 3, 5])
def test_gather_partials_gloo(world, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(29511 + world), str(script)],
                       capture_output=True, text=True, timeout=300, env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == world
