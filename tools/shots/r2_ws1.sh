#!/bin/bash
# Round 2 development shot: weights-stationary dense chain -- parity tests, chain / layer sweep, emulated-rank and N=1 quick bench lines.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dense_ws.py -m gpu -q --timeout 120 -x > gpurun_out/ws_pytest.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/ws_pytest.log
timeout 200 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_dense_ws.py -m gpu -q --timeout 150 -k "113 or (reduce and 900)" > gpurun_out/ws_memcheck.log 2>&1; echo "memcheck exit $?"; tail -5 gpurun_out/ws_memcheck.log
timeout 400 python tests/perf/ws_sweep.py > gpurun_out/ws_sweep.log 2>&1; echo "sweep exit $?"; cat gpurun_out/ws_sweep.log | tail -110
for WS in 0 1; do
  SBEV_DENSE_WS=$WS timeout 300 python bench.py --emulate-world 8 --skip-cpu --skip-backbone --skip-gpu-baseline --skip-e2e > gpurun_out/bench_emu8_ws$WS.json 2> gpurun_out/bench_emu8_ws$WS.err; echo "emu8 ws$WS exit $?"
  python -c "import json;d=json.loads(open('gpurun_out/bench_emu8_ws$WS.json').read().strip().splitlines()[-1]);print('emu8 ws$WS ms/step',d['ms_per_step'],'launches',d['launches_per_step'])"
  SBEV_DENSE_WS=$WS timeout 300 python bench.py --skip-cpu --skip-backbone --skip-gpu-baseline --skip-e2e > gpurun_out/bench_n1_ws$WS.json 2> gpurun_out/bench_n1_ws$WS.err; echo "n1 ws$WS exit $?"
  python -c "import json;d=json.loads(open('gpurun_out/bench_n1_ws$WS.json').read().strip().splitlines()[-1]);print('n1 ws$WS ms/step',d['ms_per_step'],'launches',d['launches_per_step'])"
done
