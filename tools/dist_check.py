#!/usr/bin/env python3
"""Multi-rank check of cpprob_sis_run_dist, one process per GPU (launch under torchrun, NCCL):
every rank runs the collective inference; rank 0 also runs the same inference on its own GPU alone and compares the merged
sums bit for bit.  Prints one JSON line on rank 0.  usage: torchrun --nproc-per-node N tools/dist_check.py"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import analytic  # noqa: E402
from cpprob_b200 import Engine, capi  # noqa: E402

rank, local_rank, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local_rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
g = analytic.golden()
CASES = [("gaussian_unknown_mean", [3.0, 4.0], 37 * capi.CHUNK + 4321), ("gaussian_unknown_mean", [3.0, 4.0], 5000 * capi.CHUNK + 777),
         ("gaussian_unknown_mean", [3.0, 4.0], 1000), ("linear_gaussian_1d", g["obs_linear_gaussian_32"], 9 * capi.CHUNK + 5),
         ("hmm", g["obs_hmm_64"], 11 * capi.CHUNK + 99), ("hmm", g["obs_hmm_1000"], 40_000)]
engine = Engine(device=local_rank, seed=0x5EED)
id_t = torch.zeros(capi.COMM_ID_BYTES, dtype=torch.uint8, device="cuda")
if rank == 0:
    id_t.copy_(torch.frombuffer(bytearray(capi.comm_get_id()), dtype=torch.uint8))
dist.broadcast(id_t, 0)
engine.comm_init(id_t.cpu().numpy().tobytes(), rank, world)
out = []
ok = True
for model, obs, n in CASES:
    st = engine.run_dist(model, obs, n)
    sha = hashlib.sha256(st["sums"].tobytes()).hexdigest()[:16]
    shas = [None] * world
    dist.all_gather_object(shas, sha)
    same_everywhere = len(set(shas)) == 1
    row = {"model": model, "n_obs": len(obs), "particles": n, "path": st["path"], "sha": sha, "same_on_every_rank": same_everywhere}
    if rank == 0:
        with Engine(device=local_rank, seed=0x5EED) as solo:
            ref = solo.run(model, obs, n)
        row["equals_single_gpu"] = bool((ref["sums"] == st["sums"]).all())
        ok = ok and row["equals_single_gpu"]
    ok = ok and same_everywhere
    out.append(row)
if rank == 0:
    print(json.dumps({"world": world, "ok": ok, "exchange": engine.comm_exchange(), "cases": out}))
engine.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
