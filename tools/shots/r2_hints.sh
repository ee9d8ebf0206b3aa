#!/bin/bash
# Round 2 development shot: L2 cache hints on the parameter GEMM (stores evict_first / weight loads evict_last): parity + in-situ step.
mkdir -p gpurun_out
SBEV_GEMM_L2_HINTS=3 timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 120 -x -k "gemm_split" > gpurun_out/hints_pytest.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/hints_pytest.log
run() { # name, extra args
  local name=$1; shift
  timeout 200 python bench.py --skip-cpu --skip-backbone --skip-gpu-baseline --skip-e2e "$@" > gpurun_out/bench_opt_$name.json 2> gpurun_out/bench_opt_$name.err
  python -c "import json;d=json.loads(open('gpurun_out/bench_opt_$name.json').read().strip().splitlines()[-1]);print('$name ms/step %.4f  gemm alone %.4f' % (d['ms_per_step'], d['roofline_tensor']['kernel_ms']))" 2>/dev/null || echo "$name FAILED"
}
run h0
run h1 --opt gemm_l2_hints=1
run h2 --opt gemm_l2_hints=2
run h3 --opt gemm_l2_hints=3 --timeline gpurun_out/timeline_n1_h3.json
run h0b
python - <<'PY'
import json
d=json.load(open('gpurun_out/timeline_n1_h3.json'))
print('== hints 3 step_us',d.get('step_us'), d.get('error'))
for k in d.get('kernels',[]): print('%8.2f %7.2f -> %7.2f s%s  %s'%(k['start_us'],k['dur_us'],k['start_us']+k['dur_us'],k['stream'],k['name'][:60]))
PY
