#!/bin/bash
# One GPU slot (gpurun -- 'bash tools/gpu_ab_session.sh'): whole-episode A/B timing of the step-kernel variants, the
# bench line, the launch list of the bench command and one full ncu capture of an idle and of a busy step.
# Every step has its own timeout and writes into gpurun_out/; copy what is to be kept to profiles/.
# VARIANTS / WORKLOADS / NCU_VARIANT / TAG override the defaults.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TAG=${TAG:-r2}
VARIANTS=${VARIANTS:-percharger,evl:G=1,evl:G=2,evl:G=4}
WORKLOADS=${WORKLOADS:-c3,c4,c3-1k,c2}
NCU_VARIANT=${NCU_VARIANT:-evl:G=2}
rm -f gpurun_out/steps_$TAG.log
timeout 400 python tools/ab_kernels.py --workloads $WORKLOADS --variants $VARIANTS --out gpurun_out/ab_$TAG.json > gpurun_out/ab_$TAG.log 2>&1
echo "ab rc=$?" >> gpurun_out/steps_$TAG.log
if [ -z "$SKIP_BENCH" ]; then
timeout 300 python bench.py > gpurun_out/bench_c3_$TAG.json 2> gpurun_out/bench_c3_$TAG.err
echo "bench rc=$?" >> gpurun_out/steps_$TAG.log
fi
if [ -z "$SKIP_NCU" ]; then
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum
timeout 150 ncu --metrics $M --clock-control none --csv -k regex:step_kernel --log-file gpurun_out/probe_c3_$TAG.csv python tools/ncu_probe.py --steps 40 --variants $NCU_VARIANT > gpurun_out/probe_$TAG.log 2>&1
echo "probe rc=$?" >> gpurun_out/steps_$TAG.log
timeout 150 ncu --set full --clock-control none --import-source on -k regex:evl_step_kernel -s 3 -c 1 -o gpurun_out/prof_idle_$TAG python tools/ncu_probe.py --steps 6 --variants $NCU_VARIANT > gpurun_out/prof_idle_$TAG.log 2>&1
echo "prof idle rc=$?" >> gpurun_out/steps_$TAG.log
timeout 150 ncu --set full --clock-control none --import-source on -k regex:evl_step_kernel -s 28 -c 1 -o gpurun_out/prof_busy_$TAG python tools/ncu_probe.py --steps 30 --variants $NCU_VARIANT > gpurun_out/prof_busy_$TAG.log 2>&1
echo "prof busy rc=$?" >> gpurun_out/steps_$TAG.log
fi
if [ -n "$LAUNCHES" ]; then
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --min-seconds 0.01 --steps 1 --no-cpu-baseline --skip-agent-rollout --no-extras > gpurun_out/launches_$TAG.log 2>&1
echo "launches rc=$?" >> gpurun_out/steps_$TAG.log
fi
cat gpurun_out/steps_$TAG.log
