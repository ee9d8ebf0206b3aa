#!/bin/bash
# Session-3 shot 3 (tight budget): full single-GPU suite minus the two tests that stalled shot 2, then each of those alone under a
# thread-method timeout (dumps every thread's stack and exits), then the bench (default path, then with the graph-mode public API)
mkdir -p gpurun_out
D1="tests/test_gpu_ops.py::test_autograd_deterministic_mode"
D2="tests/test_gpu_layer.py::test_layer_cuda_graph_mode_equals_eager"
timeout 150 python -m pytest tests -m gpu -q -x --timeout 60 --timeout-method=thread --deselect tests/test_gpu_multi.py --deselect $D1 --deselect $D2 > gpurun_out/c3_pytest.log 2>&1; echo "suite exit $?"; tail -4 gpurun_out/c3_pytest.log | cut -c1-300
timeout 45 python -m pytest -q -x --timeout 25 --timeout-method=thread $D1 > gpurun_out/c3_d1.log 2>&1; echo "autograd-deterministic exit $?"; tail -40 gpurun_out/c3_d1.log | cut -c1-200
timeout 45 python -m pytest -q -x --timeout 25 --timeout-method=thread $D2 > gpurun_out/c3_d2.log 2>&1; echo "graph-mode exit $?"; tail -40 gpurun_out/c3_d2.log | cut -c1-200
timeout 60 python bench.py --steps 100 --warmup 10 --skip-cpu --skip-backbone > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench exit $?"
python -c "import json;d=json.load(open('gpurun_out/bench_c3.json'));print(d['value'], d['ms_per_step'], d['launches_per_step'], d['e2e']['value'], d['e2e_resident_features'])" 2>&1 | tail -1
SBEV_BENCH_GRAPH_API=1 timeout 60 python bench.py --steps 100 --warmup 10 --skip-cpu --skip-backbone > gpurun_out/bench_c3g.json 2> gpurun_out/bench_c3g.err; echo "bench graph-api exit $?"
python -c "import json;d=json.load(open('gpurun_out/bench_c3g.json'));print(d['value'], d['ms_per_step'], d['e2e_resident_features'])" 2>&1 | tail -1
