#!/bin/bash
# GPU tests of the step kernels on the current build, then the in-session A/B of tools/sess_ab2.sh.
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
timeout 500 python -m pytest tests/test_gpu_evlist.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q --timeout 200 --timeout-method=thread > gpurun_out/test_gpu_$TAG.log 2>&1; echo "gpu tests rc=$?"; tail -2 gpurun_out/test_gpu_$TAG.log
bash tools/sess_ab2.sh
