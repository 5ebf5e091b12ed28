#!/bin/bash
# On a multi-GPU box: the bench line at N = 1 (no configs block), then N = 8/4/2 (peer exchange) and N = 8 with the NCCL
# exchange, one after the other on the same box.  usage: tools/scale_once.sh <tag> [max gpus]   -> gpurun_out/scale/<tag>_*.json
TAG=${1:-s}; MAXN=${2:-8}
D=gpurun_out/scale; mkdir -p $D
python bench.py --gpus 1 --steps 20 --warmup 3 --no-configs --cpu-particles 20000 > $D/${TAG}_1gpu.json 2> $D/${TAG}_1gpu.err
for N in 8 4 2; do
  [ $N -le $MAXN ] || continue
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2955$N bench.py --gpus $N --steps 20 --warmup 3 \
      > $D/${TAG}_${N}gpu.json 2> $D/${TAG}_${N}gpu.err
done
CPPROB_SIS_EXCHANGE=nccl python -m torch.distributed.run --nnodes=1 --nproc-per-node $MAXN --master-addr 127.0.0.1 --master-port 29559 bench.py --gpus $MAXN --steps 20 --warmup 3 \
    > $D/${TAG}_${MAXN}gpu_nccl.json 2> $D/${TAG}_${MAXN}gpu_nccl.err
python - <<PY
import json, glob
for f in sorted(glob.glob("$D/${TAG}_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], d["n_gpus"], d.get("exchange"), "weak ms", round(d["ms_per_step"], 4), "kernel ms", round(d["kernel_ms_per_step"], 4),
              "strong ms", round(d["strong_scaling"]["ms_per_step"], 4), "value %.4g" % d["value"], "e2e %.4g" % d["e2e"]["value"], d["sums_sha"]["sha256_16"], d["clocks"]["sm_mhz"] if d.get("clocks") else None)
    except Exception as ex:
        print(f, "unreadable", ex)
PY
