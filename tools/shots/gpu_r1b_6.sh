#!/bin/bash
# Session-2 shot 6: instruction-stream fixes (GEMM split epilogue, mix equal-count LN merge, SASA MUFU sqrt/exp + mask template,
# 16-byte chain staging): full single-GPU parity suite + bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x --deselect tests/test_gpu_multi.py > gpurun_out/s6_pytest.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/s6_pytest.log | cut -c1-300
timeout 300 python bench.py --steps 50 --warmup 5 --skip-cpu --breakdown > gpurun_out/bench_s6.json 2> gpurun_out/bench_s6.err
python -c "import json;d=json.load(open('gpurun_out/bench_s6.json'));b=d.get('breakdown_ms',{});print('s6', d['value'], d['ms_per_step'], {k:round(v,4) for k,v in b.items()})" 2>&1 | tail -1
