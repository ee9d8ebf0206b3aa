#!/bin/bash
# Session-2 shot 9: A-resident 128-wide parameter GEMM (gemm_impl 4): parity + A/B bench
mkdir -p gpurun_out
timeout 300 python -m pytest -q --timeout 200 -x "tests/test_gpu_ops.py::test_gemm_split_output_and_tma_fed_mix" > gpurun_out/s9_gemm.log 2>&1; echo "gemm pytest exit $?"; tail -3 gpurun_out/s9_gemm.log | cut -c1-300
for gi in 0 4; do
  SBEV_GEMM_IMPL=$gi timeout 300 python bench.py --steps 50 --warmup 5 --skip-cpu --skip-backbone > gpurun_out/bench_s9_$gi.json 2> gpurun_out/bench_s9_$gi.err
  python -c "import json;d=json.load(open('gpurun_out/bench_s9_$gi.json'));print('gemm_impl=$gi', d['value'], d['ms_per_step'], d['launches_per_step'], 'gemm alone ms', d['roofline_tensor']['kernel_ms'])" 2>&1 | tail -1
done
