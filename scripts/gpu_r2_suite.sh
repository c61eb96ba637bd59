# round 2: the full -m gpu suite + the bench
mkdir -p gpurun_out
T=${TAG:-r2s}
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/${T}_gpu_suite.log 2>&1
echo "gpu suite rc=$?"; tail -15 gpurun_out/${T}_gpu_suite.log
