"""N>1 host-side logic on CPU: world_size-2 (and 3) gloo groups shard the chunks, all-gather their
partial rows and must end up with the same, chunk-ordered matrix on every rank."""
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
