#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_all.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_all.log
timeout 600 python bench.py --steps 30 --warmup 3 --breakdown > gpurun_out/bench.json 2> gpurun_out/bench.err
SBEV_GATHER_VARIANT=1 timeout 600 python bench.py --steps 30 --warmup 3 --skip-cpu > gpurun_out/bench_v1.json 2> gpurun_out/bench_v1.err
timeout 600 python bench.py --steps 30 --warmup 3 --skip-cpu --layout nhwc > gpurun_out/bench_nhwc.json 2> gpurun_out/bench_nhwc.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-graph --skip-cpu > gpurun_out/bench_ncu.log 2>&1
tail -15 gpurun_out/pytest_all.log; tail -3 gpurun_out/bench.err
