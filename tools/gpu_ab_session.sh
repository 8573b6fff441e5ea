#!/bin/bash
# One short GPU slot: A/B timing of the step kernels, parity of the event-driven kernel, ncu counters, then the whole
# GPU suite with the event-driven kernel forced on.  Every step has its own timeout and writes into gpurun_out/.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 200 python tools/ab_kernels.py --workloads c3,c4 --steps 448 --out gpurun_out/ab_kernels.json > gpurun_out/ab.log 2>&1
echo "ab rc=$?" >> gpurun_out/steps.log
timeout 170 python -m pytest tests/test_gpu_evlist.py -x -q > gpurun_out/test_evl.log 2>&1
echo "test_evl rc=$?" >> gpurun_out/steps.log
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum
timeout 120 ncu --metrics $M --clock-control none --csv -k regex:step_kernel --log-file gpurun_out/probe_c3.csv python tools/ncu_probe.py --steps 30 > gpurun_out/probe.log 2>&1
echo "probe rc=$?" >> gpurun_out/steps.log
timeout 120 ncu --set full --clock-control none --import-source on -k regex:evl_step_kernel -s 28 -c 1 -o gpurun_out/prof_evl_g1 python tools/ncu_probe.py --steps 30 --variants evlist:1 > gpurun_out/prof.log 2>&1
echo "prof rc=$?" >> gpurun_out/steps.log
EV2B_KERNEL=evlist timeout 300 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_evlist.py > gpurun_out/test_all_evl.log 2>&1
echo "test_all_evl rc=$?" >> gpurun_out/steps.log
timeout 120 python tools/ab_kernels.py --workloads c3-1k,c2 --steps 448 --out gpurun_out/ab_kernels_small.json > gpurun_out/ab_small.log 2>&1
echo "ab_small rc=$?" >> gpurun_out/steps.log
cat gpurun_out/steps.log; tail -3 gpurun_out/ab.log gpurun_out/test_evl.log gpurun_out/test_all_evl.log
