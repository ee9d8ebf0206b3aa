#!/bin/bash
# Session-3 shot 2: deterministic backward (empty case fixed) + layer CUDA-graph mode parity; bench with graph-mode resident e2e
mkdir -p gpurun_out
timeout 300 python -m pytest -q --timeout 200 -x tests/test_gpu_ops.py tests/test_gpu_layer.py -k "deterministic or cuda_graph" > gpurun_out/c2_pytest.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/c2_pytest.log | cut -c1-400
timeout 300 python bench.py --steps 100 --warmup 10 --skip-cpu --skip-backbone > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench exit $?"; tail -3 gpurun_out/bench_c2.err
python -c "import json;d=json.load(open('gpurun_out/bench_c2.json'));print(d['value'], d['ms_per_step'], d['launches_per_step'], d['e2e'], d['e2e_resident_features'])" 2>&1 | tail -1
