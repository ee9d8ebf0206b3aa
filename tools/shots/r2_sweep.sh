#!/bin/bash
# Round 2, 1-GPU development shot: GPU tests, per-kernel sweep over shard sizes / kernel variants, chunked-mixing experiment.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log
timeout 900 python tests/perf/kernel_sweep.py "$@" > gpurun_out/kernel_sweep.log 2>&1; echo "sweep exit $?"
timeout 600 python tests/perf/mix_chunk_bench.py > gpurun_out/mix_chunk_bench.log 2>&1; echo "chunk exit $?"; cat gpurun_out/mix_chunk_bench.log | tail -8
