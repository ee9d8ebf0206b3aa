#!/bin/bash
# Session-3 shot 5 (last GPU seconds of the round): warp-batched deterministic-backward reduce -- parity first, then the whole
# single-GPU suite on the final tree, then op-level timings and the T=1 bench line (BASELINE config 2)
mkdir -p gpurun_out
timeout 60 python -m pytest -q -x --timeout 40 --timeout-method=thread tests/test_gpu_ops.py -k "deterministic or bwd" > gpurun_out/c5_det.log 2>&1; echo "det exit $?"; tail -3 gpurun_out/c5_det.log | cut -c1-300
timeout 60 python -m pytest tests -m gpu -q -x --timeout 40 --timeout-method=thread --deselect tests/test_gpu_multi.py > gpurun_out/c5_pytest.log 2>&1; echo "suite exit $?"; tail -3 gpurun_out/c5_pytest.log | cut -c1-300
timeout 45 python tests/perf/op_bench.py > gpurun_out/c5_op_bench.log 2>&1; echo "op_bench exit $?"; grep -o '"ours_bwd_ms[^}]*' gpurun_out/c5_op_bench.log | head -4
timeout 40 python bench.py --frames 1 --steps 100 --warmup 10 --skip-cpu --skip-backbone > gpurun_out/bench_c5_t1.json 2> gpurun_out/bench_c5_t1.err; echo "bench T=1 exit $?"
python -c "import json;d=json.load(open('gpurun_out/bench_c5_t1.json'));print(d['value'], d['ms_per_step'], d['launches_per_step'])" 2>&1 | tail -1
