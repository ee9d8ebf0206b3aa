#!/bin/bash
# Round 2: e2e (host buffers -> forward -> host) with torch pinned memory vs write-combined pinned memory.
mkdir -p gpurun_out
for P in default wc; do
  timeout 120 python bench.py --skip-cpu --skip-backbone --skip-gpu-baseline --pinned $P > gpurun_out/bench_pin_$P.json 2> gpurun_out/bench_pin_$P.err
  python -c "import json;d=json.loads(open('gpurun_out/bench_pin_$P.json').read().strip().splitlines()[-1]);e=d['e2e'];print('$P e2e %.1f samples/s  %.3f ms/forward  %s' % (e['value'], e['ms_per_step'], e.get('host_buffers')))" 2>/dev/null || { echo "$P FAILED"; tail -3 gpurun_out/bench_pin_$P.err; }
done
