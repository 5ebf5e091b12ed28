#!/bin/bash
# On the GPU box: bench every variant in cpprob_b200/lib/variants and read the FP64 pipe utilisation with ncu.
mkdir -p gpurun_out/sweep
for so in cpprob_b200/lib/variants/*.so; do
  n=$(basename $so .so)
  CPPROB_SIS_LIB=$PWD/$so python bench.py --steps 10 --warmup 3 --cpu-particles 1000 --no-configs --strong-particles 1000000 > gpurun_out/sweep/$n.json 2> gpurun_out/sweep/$n.err
  CPPROB_SIS_LIB=$PWD/$so ncu --metrics sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:k_sis_fused -s 1 -c 1 --csv --log-file gpurun_out/sweep/$n.ncu.csv python bench.py --steps 1 --warmup 3 --particles 200000000 --cpu-particles 1000 --no-configs --strong-particles 1000000 > /dev/null 2>&1
  python - "$n" <<'PY'
import json,sys,csv
n=sys.argv[1]
try:
    j=json.load(open(f"gpurun_out/sweep/{n}.json"))
    m={r[-3]:r[-1] for r in csv.reader(open(f"gpurun_out/sweep/{n}.ncu.csv")) if len(r)>5 and r[0].isdigit()}
    print(f"{n:8s} value={j['value']:.4e} ms/step={j['ms_per_step']:.3f} kernel_ms={j['kernel_ms_per_step']:.3f} fp64pipe={m.get('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active')} issue={m.get('smsp__issue_active.avg.pct_of_peak_sustained_active')} inst={m.get('smsp__inst_executed.sum')} ncu_ms={m.get('gpu__time_duration.sum')} mean={j['posterior']['mean']:.6f}")
except Exception as e:
    print(n, "failed", e)
PY
done
