#!/bin/bash
# Session-2 shot 5: backbone conv path timing vs stock PyTorch (cuDNN bf16 channels_last), 6 and 48 images, per-shape table
mkdir -p gpurun_out
timeout 300 python tests/perf/backbone_bench.py --images 6 --layers > gpurun_out/backbone_bench_6.json 2> gpurun_out/backbone_bench_6.err; echo "exit $?"; tail -3 gpurun_out/backbone_bench_6.err
timeout 300 python tests/perf/backbone_bench.py --images 48 > gpurun_out/backbone_bench_48.json 2> gpurun_out/backbone_bench_48.err; echo "exit $?"
python - <<'PY'
import json
for n in (6, 48):
    try:
        d = json.load(open('gpurun_out/backbone_bench_%d.json' % n))
        print(n, {k: (round(v, 4) if isinstance(v, float) else v) for k, v in d.items() if k != 'layers'})
        for r in d.get('layers', []):
            print('  cin %4d cout %4d k%d s%d %3dx%-3d x%d  ours %7.1f us %6.1f TF | cudnn %7.1f us %6.1f TF' % (r['cin'], r['cout'], r['k'], r['s'], r['h'], r['w'], r['count'], r['ours_us'], r['ours_tflops'], r['cudnn_us'], r['cudnn_tflops']))
    except Exception as e:
        print(n, 'failed', e)
PY
