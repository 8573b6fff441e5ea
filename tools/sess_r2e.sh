mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
L=gpurun_out/steps_r2e.log; rm -f $L
timeout 600 python -m pytest tests -m gpu -x -q --timeout 200 --timeout-method=thread > gpurun_out/test_gpu_r2e.log 2>&1; echo "gpu tests rc=$?" >> $L
timeout 200 python tools/ab_kernels.py --workloads c3,c4,c5 --variants evl --out gpurun_out/ab_r2e.json > gpurun_out/ab_r2e.log 2>&1; echo "ab rc=$?" >> $L
timeout 150 ncu --set full --clock-control none --import-source on -k regex:evl_step_kernel -s 28 -c 1 -o /tmp/prof_busy python tools/ncu_probe.py --steps 30 --variants evl > gpurun_out/prof_busy_r2e.log 2>&1
python tools/ncu_summary.py /tmp/prof_busy.ncu-rep > gpurun_out/r2e_evl_ncu_busy_step.txt 2>&1; echo "prof rc=$?" >> $L
timeout 150 ncu --set full --clock-control none --import-source on -k regex:evl_step_kernel -s 3 -c 1 -o /tmp/prof_idle python tools/ncu_probe.py --steps 6 --variants evl > gpurun_out/prof_idle_r2e.log 2>&1
python tools/ncu_summary.py /tmp/prof_idle.ncu-rep > gpurun_out/r2e_evl_ncu_idle_step.txt 2>&1
cp /tmp/prof_idle.ncu-rep gpurun_out/prof_idle_r2e.ncu-rep
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r2e.csv python bench.py --sweeps-only --min-seconds 0.0 --steps 1 > gpurun_out/launches_r2e.log 2>&1; echo "launches rc=$?" >> $L
cat $L; tail -5 gpurun_out/test_gpu_r2e.log; du -sh gpurun_out
