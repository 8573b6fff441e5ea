#!/bin/bash
# Per-source-line instruction / stall attribution of evl_step_kernel at a busy and at a low-occupancy step (text only).
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
TAG=${TAG:-r2g}
for spec in "busiest 37" "low 9" "idle 3"; do
  set -- $spec
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:evl_step_kernel -s $2 -c 1 -f -o /tmp/prof_$1 python tools/ncu_probe.py --steps $(($2 + 2)) --variants evl > gpurun_out/lines_$1_$TAG.log 2>&1
  python tools/ncu_summary.py /tmp/prof_$1.ncu-rep > gpurun_out/${TAG}_evl_ncu_$1_step.txt 2>&1
  python tools/ncu_lines.py /tmp/prof_$1.ncu-rep ev2gym_b200/csrc/libev2b.so evl_step_kernelIfLi2ELb1ELi1ELb0ELb0ELi128 120 > gpurun_out/${TAG}_evl_lines_$1.txt 2>&1
done
ls -la gpurun_out | tail -8; du -sh gpurun_out
