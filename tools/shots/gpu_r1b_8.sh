#!/bin/bash
# Session-2 shot 8: sample points fused into the chain epilogue (10 launches/layer): full single-GPU suite + bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x --deselect tests/test_gpu_multi.py > gpurun_out/s8_pytest.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/s8_pytest.log | cut -c1-300
for fp in 1 0; do
SBEV_DENSE_FUSE_POINTS=$fp timeout 300 python bench.py --steps 50 --warmup 5 --skip-cpu > gpurun_out/bench_s8_$fp.json 2> gpurun_out/bench_s8_$fp.err
python -c "import json;d=json.load(open('gpurun_out/bench_s8_$fp.json'));print('fuse_points=$fp', d['value'], d['ms_per_step'], d['launches_per_step'])" 2>&1 | tail -1
done
