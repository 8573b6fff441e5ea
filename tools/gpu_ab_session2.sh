#!/bin/bash
# Second short GPU slot of the session: the pipelined event-driven kernel (default from 2048 envs up).
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 150 python tools/ab_kernels.py --workloads c3,c4 --steps 448 --out gpurun_out/ab2_kernels.json > gpurun_out/ab2.log 2>&1
echo "ab2 rc=$?" >> gpurun_out/steps2.log
timeout 120 python -m pytest tests/test_gpu_evlist.py tests/test_gpu_fullsize.py -x -q > gpurun_out/test2_evl.log 2>&1
echo "test2 rc=$?" >> gpurun_out/steps2.log
timeout 170 python bench.py > gpurun_out/bench_c3_evl.json 2> gpurun_out/bench_c3_evl.err
echo "bench rc=$?" >> gpurun_out/steps2.log
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum
timeout 100 ncu --metrics $M --clock-control none --csv -k regex:step_kernel --log-file gpurun_out/probe2_c3.csv python tools/ncu_probe.py --steps 30 > gpurun_out/probe2.log 2>&1
echo "probe2 rc=$?" >> gpurun_out/steps2.log
timeout 100 ncu --set full --clock-control none --import-source on -k regex:evl_step_kernel -s 28 -c 1 -o gpurun_out/prof2_evl_g1 python tools/ncu_probe.py --steps 30 --variants evlist:1 > gpurun_out/prof2.log 2>&1
echo "prof2 rc=$?" >> gpurun_out/steps2.log
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches2.csv python bench.py --steps 64 --warmup 16 --no-cpu-baseline --skip-agent-rollout > gpurun_out/launches2.log 2>&1
echo "launches rc=$?" >> gpurun_out/steps2.log
cat gpurun_out/steps2.log
