#!/bin/bash
# Session-3 shot 1: deterministic op backward (sbev_msmv_bwd_det) parity + timing; mix work order A/B (option mix_order)
mkdir -p gpurun_out
timeout 400 python -m pytest -q --timeout 200 -x tests/test_gpu_ops.py -k "bwd or deterministic or autograd or mix or pybind" > gpurun_out/c1_pytest.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/c1_pytest.log | cut -c1-400
for mo in 0 1 0 1; do
  SBEV_MIX_ORDER=$mo timeout 300 python bench.py --steps 100 --warmup 10 --skip-cpu --skip-backbone > gpurun_out/bench_c1_$mo.json 2> gpurun_out/bench_c1_$mo.err
  python -c "import json;d=json.load(open('gpurun_out/bench_c1_$mo.json'));print('mix_order=$mo', d['value'], d['ms_per_step'], d['launches_per_step'])" 2>&1 | tail -1
done
timeout 300 python tests/perf/op_bench.py > gpurun_out/c1_op_bench.log 2>&1; echo "op_bench exit $?"; tail -5 gpurun_out/c1_op_bench.log | cut -c1-600
