mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
L=gpurun_out/steps_r2d.log; rm -f $L
timeout 600 python -m pytest tests -m gpu -x -q --timeout 200 --timeout-method=thread > gpurun_out/test_gpu_r2d.log 2>&1; echo "gpu tests rc=$?" >> $L
timeout 400 python bench.py > gpurun_out/bench_c3_r2d.json 2> gpurun_out/bench_c3_r2d.err; echo "bench rc=$?" >> $L
timeout 200 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_c3_r2d_ref.json 2> gpurun_out/bench_c3_r2d_ref.err; echo "bench ref rc=$?" >> $L
timeout 120 python tools/e2e_sweep.py > gpurun_out/e2e_sweep_r2d.jsonl 2> gpurun_out/e2e_sweep_r2d.err; echo "e2e sweep rc=$?" >> $L
for wl in c4 c5 c2; do timeout 200 python bench.py --workload $wl --no-extras --min-seconds 0.4 --no-cpu-baseline > gpurun_out/bench_${wl}_r2d.json 2> gpurun_out/bench_${wl}_r2d.err; echo "bench $wl rc=$?" >> $L; done
timeout 150 ncu --set full --clock-control none --import-source on -k regex:evl_step_kernel -s 28 -c 1 -o /tmp/prof_busy python tools/ncu_probe.py --steps 30 --variants evl > gpurun_out/prof_busy_r2d.log 2>&1
python tools/ncu_summary.py /tmp/prof_busy.ncu-rep > gpurun_out/r2d_evl_default_ncu_busy_step.txt 2>&1; echo "prof rc=$?" >> $L
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r2d.csv python bench.py --min-seconds 0.01 --steps 1 --no-cpu-baseline --skip-agent-rollout --no-extras > gpurun_out/launches_r2d.log 2>&1; echo "launches rc=$?" >> $L
cat $L; tail -5 gpurun_out/test_gpu_r2d.log; du -sh gpurun_out
