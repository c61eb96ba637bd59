mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_training.py -m gpu -q --timeout 240 2>&1 | tail -30 > gpurun_out/t_train.log
grep -E "passed|failed|^FAILED|^E  " gpurun_out/t_train.log | head
for dt in fp32 bf16; do timeout 600 python scripts/bench_train.py --dtype $dt --steps 2 --warmup 1 2>&1 | tail -3; done
timeout 600 python scripts/bench_train.py --dtype bf16 --moco --steps 2 --warmup 1 2>&1 | tail -2
timeout 300 python scripts/bench_train.py --cpu-ref 2>&1 | tail -1
