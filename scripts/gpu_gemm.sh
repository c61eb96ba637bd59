mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 120 -x -k "gemm" 2>&1 | tail -8
timeout 300 python scripts/bench_gemm.py > gpurun_out/bench_gemm.log 2>&1; cut -c1-170 gpurun_out/bench_gemm.log | tail -12
