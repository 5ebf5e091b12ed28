#!/bin/bash
# Reproduces the round-1 numbers of profiles/r01_notes.md on a B200 box (run from the repo root, after
# `python -c "import __graft_entry__ as g; g.build()"`).  Every step writes under gpurun_out/.
set -x
mkdir -p gpurun_out
python -m pytest tests -q -m gpu                                   # parity suite through the C ABI
python bench.py > gpurun_out/bench.json                            # headline: C2, 1e9 particles, one GPU
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json   # CPU arm, all host cores
bash tools/profile_gpu.sh r01z                                     # ncu launch list + full capture of k_sis_fused + row kernels
python tools/summarise_profiles.py r01z                            # -> profiles/r01z_*
python tools/bench_configs.py                                      # C3 / C4 / C5 device-timed
python tools/bench_files.py 2e7                                    # posterior-file stage, GPU text vs host text
python tools/calibration.py 128 1e9                                # z-scores of posterior mean / log-evidence over 128 seeds
# A/B of the normal sampler in the same kernel (FP64-pipe share vs throughput):
#   tools/build_variant.sh bm -DCPPROB_NORMAL_BOX_MULLER && bash tools/sweep_gpu.sh
# multi-GPU (one box): for n in 2 4 8; do python -m torch.distributed.run --nnodes=1 --nproc-per-node $n \
#   --master-addr 127.0.0.1 --master-port 29500 bench.py --gpus $n; done
