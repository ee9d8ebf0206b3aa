#!/bin/bash
# Session-2 shot 2: ncu --set full of the dense chain kernels (default and N-split) to find what they actually wait for
mkdir -p gpurun_out
for ns in 0 4; do
  SBEV_DENSE_NSPLIT=$ns timeout 280 ncu --set full --clock-control none --import-source on -k regex:"dense_chain" -s 10 -c 5 -f -o gpurun_out/prof_dense_ns$ns \
      python bench.py --steps 2 --warmup 2 --no-graph --no-overlap --skip-cpu > gpurun_out/ncu_dense_ns$ns.log 2>&1
  echo "ncu ns$ns exit $?"
done
ls -la gpurun_out/*.ncu-rep
