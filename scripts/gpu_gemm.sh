mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_training.py -m gpu -q --timeout 300 -x 2>&1 | tail -4
timeout 300 python scripts/bench_gemm.py > gpurun_out/bench_gemm.log 2>&1; cut -c1-170 gpurun_out/bench_gemm.log | tail -12
timeout 600 python scripts/bench_train.py --dtype bf16 --steps 3 --warmup 2 --graph 1 --batch 8 2>&1 | tail -1 | cut -c1-330
