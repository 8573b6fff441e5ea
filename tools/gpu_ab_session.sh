#!/bin/bash
# One short GPU slot (gpurun -- 'bash tools/gpu_ab_session.sh'): A/B timing of the step kernels on the bench workloads,
# GPU parity of the event-driven kernel, the bench line, one full ncu capture of its busy step and the launch list of
# the bench command.  Every step has its own timeout and writes into gpurun_out/; copy what is to be kept to profiles/.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 200 python tools/ab_kernels.py --workloads c3,c4,c3-1k,c2 --variants percharger:0,evlist:1,evlist:2,evlist:4,evlist:2:stage,evlist:1:stage,evlist:2:pf1,evlist:2:pf2,evlist:1:pf1 --out gpurun_out/ab_kernels.json > gpurun_out/ab.log 2>&1
echo "ab rc=$?" >> gpurun_out/steps.log
timeout 120 python -m pytest tests/test_gpu_evlist.py tests/test_gpu_fullsize.py -x -q > gpurun_out/test_evl.log 2>&1
echo "test_evl rc=$?" >> gpurun_out/steps.log
timeout 120 python bench.py > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
echo "bench rc=$?" >> gpurun_out/steps.log
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum
timeout 100 ncu --metrics $M --clock-control none --csv -k regex:step_kernel --log-file gpurun_out/probe_c3.csv python tools/ncu_probe.py --steps 30 > gpurun_out/probe.log 2>&1
echo "probe rc=$?" >> gpurun_out/steps.log
timeout 100 ncu --set full --clock-control none --import-source on -k regex:evl_step_kernel -s 28 -c 1 -o gpurun_out/prof_evl python tools/ncu_probe.py --steps 30 --variants evlist:2 > gpurun_out/prof.log 2>&1
echo "prof rc=$?" >> gpurun_out/steps.log
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 64 --warmup 16 --no-cpu-baseline --skip-agent-rollout > gpurun_out/launches.log 2>&1
echo "launches rc=$?" >> gpurun_out/steps.log
cat gpurun_out/steps.log
