mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_training.py -m gpu -q --timeout 300 2>&1 | tail -60 > gpurun_out/t_all.log
grep -E "passed|failed|^FAILED|^E  " gpurun_out/t_all.log | head -30
for dt in fp32 bf16; do timeout 600 python scripts/bench_train.py --dtype $dt --steps 3 --warmup 1 2>&1 | tail -1; done
timeout 600 python scripts/bench_train.py --dtype bf16 --moco --steps 3 --warmup 1 2>&1 | tail -1
