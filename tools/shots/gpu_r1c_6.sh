#!/bin/bash
# Session-3 shot 6 (the round's last GPU seconds): single-GPU suite on the final tree
mkdir -p gpurun_out
timeout 28 python -m pytest tests -m gpu -q -x --timeout 20 --timeout-method=thread --deselect tests/test_gpu_multi.py > gpurun_out/c6_pytest.log 2>&1; echo "suite exit $?"; tail -3 gpurun_out/c6_pytest.log | cut -c1-300
