#!/bin/bash
# Last slot: the bench line, one full ncu capture and the launch list of the final build.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 75 python bench.py > gpurun_out/bench5_c3.json 2> gpurun_out/bench5_c3.err
echo "bench rc=$?" >> gpurun_out/steps5.log
timeout 40 ncu --set full --clock-control none --import-source on -k regex:evl_step_kernel -s 28 -c 1 -o gpurun_out/prof5_evl_g2 python tools/ncu_probe.py --steps 30 --variants evlist:2 > gpurun_out/prof5.log 2>&1
echo "prof5 rc=$?" >> gpurun_out/steps5.log
timeout 45 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches5.csv python bench.py --steps 64 --warmup 16 --no-cpu-baseline --skip-agent-rollout > gpurun_out/launches5.log 2>&1
echo "launches rc=$?" >> gpurun_out/steps5.log
cat gpurun_out/steps5.log
