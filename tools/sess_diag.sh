mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
timeout 400 python -m pytest tests -m gpu -x -v --timeout 90 --timeout-method=thread > gpurun_out/test_gpu_diag.log 2>&1; echo "gpu tests rc=$?" > gpurun_out/steps_diag.log
for v in percharger evl:G=4 evl:G=1; do
timeout 60 python tools/ab_kernels.py --workloads c3 --variants $v --min-seconds 0.05 > gpurun_out/ab_diag_$v.log 2>&1; echo "ab $v rc=$?" >> gpurun_out/steps_diag.log
done
cat gpurun_out/steps_diag.log; tail -30 gpurun_out/test_gpu_diag.log
