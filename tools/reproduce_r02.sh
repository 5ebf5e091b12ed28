#!/bin/bash
# Reproduces the round-2 measurements.  Run from the repo root in the build container; every GPU step goes through gpurun.
# (The numbers in DESIGN.md / README.md / profiles/r02_notes.md were taken with exactly these commands.)
set -e
python -c "import __graft_entry__ as g; g.build()"                       # nvcc sm_100a + oracle + oracle/_ref + examples
python -m pytest tests -q -m "not gpu"                                    # oracle vs the reference's own code and pins, host logic
G=/usr/local/graft/bin/gpurun
$G --timeout 1500 -- 'python -m pytest tests -m gpu -q | tail -3; python bench.py > gpurun_out/r02_bench.json; python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_ref.json'
$G --timeout 1800 -- 'bash tools/profile_r02.sh r02'                      # ncu: launch list, per-kernel metrics, full captures (CSV pages)
python tools/summarise_r02.py r02                                         # -> profiles/r02_*.csv, profiles/kernel_costs.json
$G --timeout 900 -- 'python tools/calibration_staged.py 64 22; python tools/fixed_overhead.py'
$G --timeout 2400 -- 'for t in memcheck racecheck initcheck; do compute-sanitizer --tool $t python tools/sanitize_once.py | tail -3; done'
bash tools/build_variant.sh bm -DCPPROB_NORMAL_BOX_MULLER && bash tools/build_variant.sh zig
$G --timeout 1200 -- 'bash tools/sweep_gpu.sh'                            # Box-Muller vs ziggurat A/B (profiles/r02_bm_ab/)
# multi-GPU (charged N x): tests of the library's exchange (peer memory and NCCL) and of multi-GPU emission, memcheck over the
# peer exchange, then every scaling point on ONE box (N = 1, 8, 4, 2 with the peer exchange, N = 8 with NCCL)
$G --gpus 2 --timeout 1200 -- 'python -m pytest tests/test_dist_gpu.py tests/test_files_gpu.py -q | tail -2; compute-sanitizer --tool memcheck python tools/sanitize_once.py | tail -3'
$G --gpus 8 --timeout 900 -- 'bash tools/scale_once.sh r02z 8'          # -> gpurun_out/scale/r02z_*.json (copied to profiles/r02z_scale_*.json)
# second half of the round: fixed cost of an inference, final build
$G --timeout 900 -- 'python tools/strong_breakdown.py 1.25e8 30 sha; python tools/strong_breakdown.py 1e9 15; python tools/strong_breakdown.py 1e4 40'
$G --timeout 2400 -- 'python -m pytest tests -m gpu -q | tail -3; python bench.py > gpurun_out/r02z_bench.json; python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02z_ref.json; bash tools/profile_r02.sh r02z'
python tools/summarise_r02.py r02z
