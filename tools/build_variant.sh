#!/bin/bash
# Builds a kernel-variant copy of libcpprob_sis.so for A/B measurements on the GPU box:
#   tools/build_variant.sh <name> [-DCPPROB_...=...]...   ->  cpprob_b200/lib/variants/<name>.so
# Select it with CPPROB_SIS_LIB=cpprob_b200/lib/variants/<name>.so python bench.py ...
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
NAME=$1; shift
OUT=$ROOT/cpprob_b200/lib/variants
mkdir -p $OUT/obj_$NAME
FLAGS="-std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -I$ROOT/include -I$ROOT/cpprob_b200/csrc --expt-relaxed-constexpr $*"
for f in sis_capi models_builtin; do
  nvcc $FLAGS -Xptxas -v -c $ROOT/cpprob_b200/csrc/$f.cu -o $OUT/obj_$NAME/$f.o 2> $OUT/obj_$NAME/$f.log &
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $OUT/$NAME.so $OUT/obj_$NAME/sis_capi.o $OUT/obj_$NAME/models_builtin.o -lcudart
echo "$NAME: $(grep -A2 'k_sis_fusedIN6models27gaussian_unknown_mean_modelELi1' $OUT/obj_$NAME/models_builtin.log | grep -E 'Used' | sed 's/ptxas info    ://')"
