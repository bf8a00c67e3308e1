#!/bin/bash
# Experiment builds of libb2az.so: tools/build_variants.sh name:"-DFLAG=.. -DFLAG2=.." ...  -> build/variants/<name>.so
# (measured on the GPU box with tools/variant_bench.py build/variants/*.so)
ROOT=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p "$ROOT/build/variants"
FL="-std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -shared -diag-suppress 177"
pids=()
for v in "$@"; do
  n=${v%%:*}; f=${v#*:}
  /usr/local/cuda/bin/nvcc $FL $f "$ROOT/alphazero-pybind11_b200/csrc/az_engine.cu" -o "$ROOT/build/variants/$n.so" &
  pids+=($!)
done
rc=0
for p in "${pids[@]}"; do wait $p || rc=1; done
ls -la "$ROOT/build/variants/"
exit $rc
