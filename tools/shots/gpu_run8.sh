#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_all.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_all.log
timeout 600 python bench.py --steps 50 --warmup 5 --breakdown > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 python bench.py --steps 50 --warmup 5 --skip-cpu --no-overlap > gpurun_out/bench_nooverlap.json 2> gpurun_out/bench_nooverlap.err
timeout 600 python bench.py --steps 50 --warmup 5 --skip-cpu --layout nhwc > gpurun_out/bench_nhwc.json 2> gpurun_out/bench_nhwc.err
timeout 600 python bench.py --steps 50 --warmup 5 --skip-cpu --precision bf16 > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err
tail -8 gpurun_out/pytest_all.log; tail -3 gpurun_out/bench.err
for f in bench bench_nooverlap bench_nhwc bench_bf16; do python -c "import json;d=json.load(open('gpurun_out/$f.json'));print('$f',d['value'],d['ms_per_step'],d['e2e_resident_features']['value'])"; done
