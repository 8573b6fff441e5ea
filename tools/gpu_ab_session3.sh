#!/bin/bash
# Third short GPU slot: A/B of the two EV-loop builds (plain vs register-prefetched), then the bench line, one full ncu
# capture and the launch list with the faster one.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
L=ev2gym_b200/csrc
timeout 90 python tools/ab_kernels.py --workloads c3,c4,c3-1k --out gpurun_out/ab3_plain.json > gpurun_out/ab3_plain.log 2>&1
echo "ab3 plain rc=$?" >> gpurun_out/steps3.log
EV2B_LIB=$PWD/$L/libev2b_pipe.so timeout 70 python tools/ab_kernels.py --workloads c3,c4 --variants evlist:1,evlist:2,evlist:4 --out gpurun_out/ab3_pipe.json > gpurun_out/ab3_pipe.log 2>&1
echo "ab3 pipe rc=$?" >> gpurun_out/steps3.log
WIN=$(python - <<'PY'
import json
def us(f):
    for l in open(f):
        d = json.loads(l)
        if d.get("workload") == "c3" and d.get("G") == 2 and "us_per_launch" in d: return d["us_per_launch"]
    return 1e9
try:
    print("pipe" if us("gpurun_out/ab3_pipe.json") < us("gpurun_out/ab3_plain.json") else "plain")
except Exception:
    print("plain")
PY
)
echo "winner=$WIN" >> gpurun_out/steps3.log
if [ "$WIN" = "pipe" ]; then export EV2B_LIB=$PWD/$L/libev2b_pipe.so; fi
timeout 120 python bench.py > gpurun_out/bench3_c3.json 2> gpurun_out/bench3_c3.err
echo "bench rc=$?" >> gpurun_out/steps3.log
timeout 80 ncu --set full --clock-control none --import-source on -k regex:evl_step_kernel -s 28 -c 1 -o gpurun_out/prof3_evl_g2 python tools/ncu_probe.py --steps 30 --variants evlist:2 > gpurun_out/prof3.log 2>&1
echo "prof3 rc=$?" >> gpurun_out/steps3.log
timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches3.csv python bench.py --steps 64 --warmup 16 --no-cpu-baseline --skip-agent-rollout > gpurun_out/launches3.log 2>&1
echo "launches rc=$?" >> gpurun_out/steps3.log
cat gpurun_out/steps3.log
