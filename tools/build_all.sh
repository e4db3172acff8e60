#!/bin/bash
# Build the library, the oracle and the GPU microbenchmarks (run from anywhere).
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
cd "$ROOT"
python -c "import __graft_entry__ as g; g.build()"
cd "$ROOT/tools/microbench"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -DFGP_HEAD_TIMING -o head_phases head_phases.cu
[ -f fp64_latency ] || nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o fp64_latency fp64_latency.cu
echo "build ok"
