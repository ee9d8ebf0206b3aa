#!/bin/bash
# Session-2 shot 7: gather L2-prefetch variant (parity + A/B), ncu launch list of the current default
mkdir -p gpurun_out
SBEV_GATHER_VARIANT=3 timeout 400 python -m pytest -q --timeout 300 -x tests/test_gpu_layer.py tests/test_gpu_ops.py -k "sampling or layer or decoder or fused" > gpurun_out/s7_gv3.log 2>&1; echo "gv3 pytest exit $?"; tail -3 gpurun_out/s7_gv3.log | cut -c1-300
for gv in 2 3; do
  SBEV_GATHER_VARIANT=$gv timeout 300 python bench.py --steps 50 --warmup 5 --skip-cpu > gpurun_out/bench_gv$gv.json 2> gpurun_out/bench_gv$gv.err
  python -c "import json;d=json.load(open('gpurun_out/bench_gv$gv.json'));print('gv$gv', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline_tensor']['kernel_ms'])" 2>&1 | tail -1
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_s7.csv python bench.py --steps 2 --warmup 2 --no-graph --no-overlap --skip-cpu > gpurun_out/bench_ncu_s7.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/launches_s7.csv')) if len(r) > 5]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
agg = collections.defaultdict(list)
for r in rows[1:]:
    try: agg[r[ki][:70]].append(float(r[vi].replace(',', '')))
    except ValueError: pass
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print('%-72s n=%3d avg %8.2f us' % (k, len(v), sum(v) / len(v) / 1000.0 if max(v) > 1000 else sum(v) / len(v)))
PY
