#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_all.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_all.log
timeout 600 python bench.py --steps 50 --warmup 5 --breakdown > gpurun_out/bench.json 2> gpurun_out/bench.err
SBEV_GATHER_VARIANT=2 timeout 600 python bench.py --steps 50 --warmup 5 --skip-cpu > gpurun_out/bench_g2.json 2> gpurun_out/bench_g2.err
SBEV_GATHER_VARIANT=0 timeout 600 python bench.py --steps 50 --warmup 5 --skip-cpu > gpurun_out/bench_g0.json 2> gpurun_out/bench_g0.err
tail -8 gpurun_out/pytest_all.log; tail -3 gpurun_out/bench.err
for f in bench bench_g2 bench_g0; do python -c "import json;d=json.load(open('gpurun_out/$f.json'));print('$f',d['value'],d['ms_per_step'],d['roofline']['kernel_ms'],d['launches_per_step'])"; done
