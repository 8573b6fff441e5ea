#!/bin/bash
# Fourth short GPU slot: previous build vs the in-place list / speculative first entry build, parity of the new one.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
L=$PWD/ev2gym_b200/csrc
timeout 60 python tools/ab_kernels.py --workloads c3,c4 --variants percharger:0,evlist:2,evlist:1 --out gpurun_out/ab4_new.json > gpurun_out/ab4_new.log 2>&1
echo "ab4 new rc=$?" >> gpurun_out/steps4.log
EV2B_LIB=$L/libev2b_prev.so timeout 50 python tools/ab_kernels.py --workloads c3,c4 --variants evlist:2,evlist:1 --out gpurun_out/ab4_prev.json > gpurun_out/ab4_prev.log 2>&1
echo "ab4 prev rc=$?" >> gpurun_out/steps4.log
timeout 60 python -m pytest tests/test_gpu_evlist.py tests/test_gpu_fullsize.py -x -q > gpurun_out/test4.log 2>&1
echo "test4 rc=$?" >> gpurun_out/steps4.log
cat gpurun_out/steps4.log; tail -2 gpurun_out/test4.log
