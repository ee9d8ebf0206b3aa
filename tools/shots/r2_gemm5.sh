#!/bin/bash
# Round 2 development shot: 3-stage CTA-pair parameter GEMM (gemm_impl 5): parity, isolated time, in-situ timeline of the N=1 step.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 120 -x -k "gemm" > gpurun_out/gemm5_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/gemm5_pytest.log
for IMPL in 0 5 2; do
  SBEV_GEMM_IMPL=$IMPL timeout 300 python tests/perf/kernel_sweep.py "param_gemm M" > gpurun_out/sweep_gemm$IMPL.log 2>&1; echo "sweep impl $IMPL exit $?"; grep param_gemm gpurun_out/sweep_gemm$IMPL.log
  SBEV_GEMM_IMPL=$IMPL timeout 300 python bench.py --skip-cpu --skip-backbone --skip-gpu-baseline --skip-e2e --timeline gpurun_out/timeline_n1_g$IMPL.json > gpurun_out/bench_n1_g$IMPL.json 2> gpurun_out/bench_n1_g$IMPL.err; echo "n1 impl $IMPL exit $?"
  SBEV_GEMM_IMPL=$IMPL timeout 300 python bench.py --no-overlap --skip-cpu --skip-backbone --skip-gpu-baseline --skip-e2e > gpurun_out/bench_n1_noov_g$IMPL.json 2> gpurun_out/bench_n1_noov_g$IMPL.err; echo "n1 no-overlap impl $IMPL exit $?"
done
python - <<'PY'
import json
for i in (0,5,2):
    for f in ('bench_n1_g%d'%i,'bench_n1_noov_g%d'%i):
        try:
            d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f,'ms/step',d['ms_per_step'], 'gemm roof', (d.get('roofline_tensor') or {}).get('kernel_ms'))
        except Exception as e: print(f,'ERR',e)
    try:
        d=json.load(open('gpurun_out/timeline_n1_g%d.json'%i))
        print('== impl',i,'step_us',d.get('step_us'), d.get('error'))
        for k in d.get('kernels',[]): print('%8.2f %7.2f -> %7.2f s%s  %s'%(k['start_us'],k['dur_us'],k['start_us']+k['dur_us'],k['stream'],k['name'][:60]))
    except Exception as e: print('timeline',i,'ERR',e)
PY
