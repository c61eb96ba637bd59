mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_training.py -m gpu -q --timeout 300 -x 2>&1 | tail -15
for g in 1; do timeout 600 python scripts/bench_train.py --dtype bf16 --steps 4 --warmup 2 --graph $g 2>&1 | tail -1 | cut -c1-400; done
timeout 600 python scripts/bench_train.py --dtype bf16 --steps 1 --warmup 2 --graph 0 --profile > gpurun_out/train_profile_bf16.log 2>&1; head -36 gpurun_out/train_profile_bf16.log | cut -c1-90,150-230
