#!/bin/bash
# Round 2: decoder-level CUDA graph -- graph-mode tests + the default bench line (forward_resident_features record).
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_layer.py -m gpu -q --timeout 200 -k "graph" > gpurun_out/pytest_graph.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_graph.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"; tail -2 gpurun_out/bench_n1.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'resident', d['e2e_resident_features'].get('ms_per_step'))
print('forward_resident', json.dumps(d.get('forward_resident_features')))
PY
