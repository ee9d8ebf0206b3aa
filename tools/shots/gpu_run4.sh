#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_all.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_all.log
timeout 600 python bench.py --steps 30 --warmup 3 --breakdown > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 python tests/perf/ref_layer_gpu.py > gpurun_out/ref_layer_gpu.log 2>&1
tail -15 gpurun_out/pytest_all.log; tail -3 gpurun_out/bench.err; tail -3 gpurun_out/ref_layer_gpu.log
