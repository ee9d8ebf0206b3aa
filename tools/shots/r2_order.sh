#!/bin/bash
# Round 2 development shot: order of the gather and the parameter GEMM (they do not overlap: both near the L2 throughput cap) x mix work order.
mkdir -p gpurun_out
run() { # name, env..., -- extra args
  local name=$1; shift
  timeout 200 env "$@" python bench.py --skip-cpu --skip-backbone --skip-gpu-baseline --skip-e2e --timeline gpurun_out/timeline_$name.json > gpurun_out/bench_ord_$name.json 2> gpurun_out/bench_ord_$name.err
  python -c "import json;d=json.loads(open('gpurun_out/bench_ord_$name.json').read().strip().splitlines()[-1]);print('$name ms/step %.4f' % d['ms_per_step'])" 2>/dev/null || echo "$name FAILED"
}
run o0 SBEV_PHASE_ORDER=0
run o1 SBEV_PHASE_ORDER=1
run o1m1 SBEV_PHASE_ORDER=1 SBEV_MIX_ORDER=1
run o2 SBEV_PHASE_ORDER=2
run o2m1 SBEV_PHASE_ORDER=2 SBEV_MIX_ORDER=1
run o0m1 SBEV_PHASE_ORDER=0 SBEV_MIX_ORDER=1
python - <<'PY'
import json
for n in ('o1m1','o2m1'):
    d=json.load(open('gpurun_out/timeline_%s.json'%n))
    print('==',n,'step_us',d.get('step_us'), d.get('error'))
    for k in d.get('kernels',[]): print('%8.2f %7.2f -> %7.2f s%s  %s'%(k['start_us'],k['dur_us'],k['start_us']+k['dur_us'],k['stream'],k['name'][:60]))
PY
