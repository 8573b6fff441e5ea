#!/bin/bash
# Quick A/B slot: GPU tests of the step kernels + whole-episode timing of the default kernel selection.
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
TAG=${TAG:-ab}
timeout 500 python -m pytest tests/test_gpu_evlist.py tests/test_gpu_parity.py -m gpu -x -q --timeout 200 --timeout-method=thread > gpurun_out/test_gpu_$TAG.log 2>&1; echo "gpu tests rc=$?"
timeout 300 python tools/ab_kernels.py --workloads ${WL:-c3,c4,c5} --variants ${VARIANTS:-evl} --out gpurun_out/ab_$TAG.json > gpurun_out/ab_$TAG.log 2>&1; echo "ab rc=$?"
tail -3 gpurun_out/test_gpu_$TAG.log; cat gpurun_out/ab_$TAG.json | head -50
