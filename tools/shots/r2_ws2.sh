#!/bin/bash
# Round 2 development shot: phase timeline + source-level ncu capture of the weights-stationary chain kernel.
mkdir -p gpurun_out
timeout 300 python tests/perf/ws_timeline.py 113 900 > gpurun_out/ws_timeline.log 2>&1; echo "timeline exit $?"; cat gpurun_out/ws_timeline.log | tail -40
timeout 400 ncu --set full --import-source on -k regex:dense_chain_ws -s 4 -c 1 -f -o gpurun_out/ws_cls113 python tests/perf/ws_sweep.py "D_cls M113 ws" > gpurun_out/ncu_ws.log 2>&1; echo "ncu exit $?"
tail -3 gpurun_out/ncu_ws.log
