#!/bin/bash
# Session-2 shot 1: N-split cluster dense chain -- parity (each variant in its own process: a trap must not poison the rest) + bench A/B
mkdir -p gpurun_out
T=tests/test_gpu_ops.py
run() { name=$1; shift; timeout 240 python -m pytest -q -x --timeout 200 "$@" > gpurun_out/ns_$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/ns_summary.log; tail -3 gpurun_out/ns_$name.log; }
run ops2 "$T::test_dense_chain_vs_torch[12]" "$T::test_dense_chain_with_fused_splitk_reduce[12-256-18-True]" "$T::test_dense_chain_with_fused_splitk_reduce[12-256-1-False]"
run ops4 "$T::test_dense_chain_vs_torch[14]" "$T::test_dense_chain_with_fused_splitk_reduce[14-256-18-True]" "$T::test_dense_chain_with_fused_splitk_reduce[14-128-5-True]"
run layer2 "tests/test_gpu_layer.py::test_decoder_layer_with_nsplit_cluster_chain[2]"
run layer4 "tests/test_gpu_layer.py::test_decoder_layer_with_nsplit_cluster_chain[4]"
for ns in 0 2 4; do
  SBEV_DENSE_NSPLIT=$ns timeout 300 python bench.py --steps 40 --warmup 5 --skip-cpu --breakdown > gpurun_out/bench_ns$ns.json 2> gpurun_out/bench_ns$ns.err
  cp gpurun_out/breakdown.json gpurun_out/breakdown_ns$ns.json 2>/dev/null
  python -c "import json;d=json.load(open('gpurun_out/bench_ns$ns.json'));print('ns$ns', d['value'], d['ms_per_step'], d.get('breakdown_ms'))" 2>&1 | tail -1
done
