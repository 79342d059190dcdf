#!/bin/bash
# build_s3d.sh NAME [nvcc -D flags...] : builds build/bin/s3d_NAME from tools/s3d_bench.cu (kernel experiments)
name=$1; shift
cd "$(dirname "$0")/.." && mkdir -p build/bin && \
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr \
  -Xptxas -v -I lua-multigrid-poisson_b200/csrc "$@" -o build/bin/s3d_$name tools/s3d_bench.cu -lcuda 2> /tmp/s3d_$name.log
rc=$?
grep -A2 "k_stream3d" /tmp/s3d_$name.log | grep -E "registers|spill" | paste - - | sed 's/ptxas info    ://g' 
echo "$name rc=$rc"
