#!/bin/bash
# Session-2 shot 3: vectorised chain epilogue (parity + A/B bench), backbone conv path parity (own process)
mkdir -p gpurun_out
timeout 400 python -m pytest -q --timeout 300 tests/test_gpu_ops.py -k "dense" tests/test_gpu_layer.py > gpurun_out/s3_dense.log 2>&1; echo "dense+layer exit $?"; tail -4 gpurun_out/s3_dense.log
timeout 400 python -m pytest -q --timeout 300 tests/test_gpu_backbone.py > gpurun_out/s3_backbone.log 2>&1; echo "backbone exit $?"; tail -25 gpurun_out/s3_backbone.log | cut -c1-400
for v in 0 1; do
  SBEV_DENSE_VEC4=$v timeout 300 python bench.py --steps 40 --warmup 5 --skip-cpu --breakdown > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err
  python -c "import json;d=json.load(open('gpurun_out/bench_v$v.json'));b=d.get('breakdown_ms',{});print('vec4=$v', d['value'], d['ms_per_step'], {k:round(b[k],4) for k in ('pos_enc_chain','ffn_cls_chain','reg_refine_chain','sasa.in_proj_tau') if k in b})" 2>&1 | tail -1
done
