#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_all.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_all.log
timeout 600 python bench.py --steps 30 --warmup 3 --breakdown > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"dense_chain|mix_mma|sasa_mma|persistent" -s 13 -c 13 -o gpurun_out/prof_layer \
    python bench.py --steps 2 --warmup 1 --no-graph --skip-cpu > gpurun_out/bench_ncu2.log 2>&1
tail -15 gpurun_out/pytest_all.log; tail -3 gpurun_out/bench.err
