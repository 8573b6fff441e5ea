#!/bin/bash
# One GPU slot: `gpurun --timeout 2400 -- 'TAG=r2f bash tools/gpu_session.sh'`.
#   1. the GPU test suite (per-test timeout: a hung kernel must not eat the slot)
#   2. whole-episode A/B of the default kernel selection on c3 / c4 / c5 (tools/ab_kernels.py)
#   3. bench.py (c3, full line incl. extras and CPU baselines) and the other workloads (short)
#   4. ncu --set full of the busiest and of an idle step (summaries only; the reports stay on the box) and the DRAM traffic
#      of the busiest step stamped with the hash of the sources -> gpurun_out/roofline_traffic_$TAG.json
#   5. the launch list of the timed sweeps (ncu --metrics gpu__time_duration.sum)
# Everything lands in gpurun_out/ (keep it far below 64 MiB: larger directories are not copied back).
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
TAG=${TAG:-r2}
L=gpurun_out/steps_$TAG.log; rm -f $L
timeout 600 python -m pytest tests -m gpu -x -q --timeout 200 --timeout-method=thread > gpurun_out/test_gpu_$TAG.log 2>&1; echo "gpu tests rc=$?" >> $L
timeout 200 python tools/ab_kernels.py --workloads c3,c4,c5 --variants ${VARIANTS:-evl,evl:G=2} --out gpurun_out/ab_$TAG.json > gpurun_out/ab_$TAG.log 2>&1; echo "ab rc=$?" >> $L
timeout 500 python bench.py > gpurun_out/bench_c3_$TAG.json 2> gpurun_out/bench_c3_$TAG.err; echo "bench rc=$?" >> $L
for wl in c4 c5 c2; do timeout 200 python bench.py --workload $wl --no-extras --min-seconds 0.4 --no-cpu-baseline > gpurun_out/bench_${wl}_$TAG.json 2> gpurun_out/bench_${wl}_$TAG.err; echo "bench $wl rc=$?" >> $L; done
timeout 150 ncu --set full --clock-control none --import-source on -k regex:evl_step_kernel -s 37 -c 1 -o /tmp/prof_busy python tools/ncu_probe.py --steps 39 --variants evl > gpurun_out/prof_busy_$TAG.log 2>&1
python tools/ncu_summary.py /tmp/prof_busy.ncu-rep > gpurun_out/${TAG}_evl_ncu_busiest_step.txt 2>&1; echo "prof busy rc=$?" >> $L
python - <<PY > gpurun_out/roofline_traffic_$TAG.json
import json, re, sys
sys.path.insert(0, ".")
from bench import source_sha
txt = open("gpurun_out/${TAG}_evl_ncu_busiest_step.txt").read()
def val(name):
    m = re.search(name + r"\s+([0-9.]+) (\w+)", txt)
    return float(m.group(1)) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[m.group(2)]
b = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
print(json.dumps({"c3": {"dram_bytes_per_launch": b, "source_sha": source_sha(),
      "source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of ev2b::evl_step_kernel<float,2,true,1,false,false> at episode step 37 of 112 (the busiest: 130 of 200 ports occupied), c3, profiles/${TAG}_evl_ncu_busiest_step.txt"}}, indent=1))
PY
timeout 150 ncu --set full --clock-control none --import-source on -k regex:evl_step_kernel -s 3 -c 1 -o /tmp/prof_idle python tools/ncu_probe.py --steps 6 --variants evl > gpurun_out/prof_idle_$TAG.log 2>&1
python tools/ncu_summary.py /tmp/prof_idle.ncu-rep > gpurun_out/${TAG}_evl_ncu_idle_step.txt 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2200 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --sweeps-only --min-seconds 0.0 --steps 1 > gpurun_out/launches_$TAG.log 2>&1; echo "launches rc=$?" >> $L
cat $L; tail -5 gpurun_out/test_gpu_$TAG.log; du -sh gpurun_out
