#!/bin/bash
# Round 2 development shot: co-residency of the fused gather next to a persistent kernel that holds S bytes of shared memory per SM.
mkdir -p gpurun_out
timeout 300 python tests/perf/coresidency.py > gpurun_out/coresidency.log 2>&1; echo "exit $?"; tail -20 gpurun_out/coresidency.log
