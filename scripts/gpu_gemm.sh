mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 300 -x -k "gemm or linear or conv1d" 2>&1 | tail -5
timeout 300 python scripts/bench_gemm.py > gpurun_out/bench_gemm.log 2>&1; cut -c1-170 gpurun_out/bench_gemm.log
