#!/bin/bash
# First GPU pass: parity tests, op bench vs the reference CUDA kernel, decoder-layer bench, ncu launch list + full capture.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 -x --deselect tests/test_gpu_layer.py > gpurun_out/pytest_ops.log 2>&1
echo "pytest ops exit $?" >> gpurun_out/pytest_ops.log
timeout 900 python -m pytest tests/test_gpu_layer.py -m gpu -q --timeout 300 > gpurun_out/pytest_layer.log 2>&1
echo "pytest layer exit $?" >> gpurun_out/pytest_layer.log
timeout 600 python tests/perf/op_bench.py > gpurun_out/op_bench.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 --breakdown > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-graph --skip-cpu > gpurun_out/bench_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sampling4d -s 2 -c 2 -o gpurun_out/prof_gather \
    python bench.py --steps 2 --warmup 1 --no-graph --skip-cpu > gpurun_out/bench_ncu2.log 2>&1
tail -5 gpurun_out/pytest_ops.log gpurun_out/pytest_layer.log gpurun_out/op_bench.log gpurun_out/bench.json gpurun_out/bench.err
