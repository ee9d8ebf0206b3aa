#!/bin/bash
# Session-2 shot 4: ncu --set full of every kernel of one decoder layer (source-level stall attribution) + launch list
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"dense_chain|mix_|sasa_v3|persistent|sampling4d|sample_points" -s 22 -c 11 -f -o gpurun_out/prof_layer_s2 \
    python bench.py --steps 2 --warmup 2 --no-graph --no-overlap --skip-cpu > gpurun_out/ncu_layer_s2.log 2>&1
echo "ncu exit $?"; ls -la gpurun_out/prof_layer_s2.ncu-rep
