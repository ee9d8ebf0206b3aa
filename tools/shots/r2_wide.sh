#!/bin/bash
# Round 2 development shot: 16-row CTAs for cls || reg -- parity, N=1 step with / without, timelines.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_dense_ws.py -m gpu -q --timeout 120 -x -k "wide_cta" > gpurun_out/wide_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/wide_pytest.log
for W in 1 0; do
  SBEV_WIDE_CTA_HEADS=$W timeout 300 python bench.py --skip-cpu --skip-backbone --skip-gpu-baseline --skip-e2e --timeline gpurun_out/timeline_n1_w$W.json > gpurun_out/bench_n1_w$W.json 2> gpurun_out/bench_n1_w$W.err; echo "n1 wide $W exit $?"
done
python - <<'PY'
import json
for i in (1,0):
    try:
        d=json.loads(open('gpurun_out/bench_n1_w%d.json'%i).read().strip().splitlines()[-1]); print('wide',i,'ms/step',d['ms_per_step'])
    except Exception as e: print('ERR',e)
    try:
        d=json.load(open('gpurun_out/timeline_n1_w%d.json'%i))
        print('== wide',i,'step_us',d.get('step_us'), d.get('error'))
        for k in d.get('kernels',[])[-5:]: print('%8.2f %7.2f -> %7.2f s%s  %s'%(k['start_us'],k['dur_us'],k['start_us']+k['dur_us'],k['stream'],k['name'][:60]))
    except Exception as e: print('timeline',i,'ERR',e)
PY
timeout 600 python -m pytest tests/test_gpu_layer.py tests/test_gpu_ops.py -m gpu -q --timeout 300 -x > gpurun_out/wide_pytest2.log 2>&1; echo "pytest layer+ops exit $?"; tail -3 gpurun_out/wide_pytest2.log
