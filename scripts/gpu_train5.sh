timeout 600 python -m pytest tests/test_gpu_training.py tests/test_gpu_model.py -m gpu -q --timeout 300 -x 2>&1 | tail -3
for b in 8 64; do timeout 600 python scripts/bench_train.py --dtype bf16 --steps 3 --warmup 2 --graph 1 --batch $b 2>&1 | tail -1 | cut -c1-330; done
timeout 600 python scripts/bench_train.py --dtype bf16 --steps 3 --warmup 2 --graph 1 --config gpt2 --batch 32 2>&1 | tail -1 | cut -c1-330
timeout 600 python scripts/bench_train.py --dtype bf16 --steps 3 --warmup 2 --graph 1 --moco --batch 64 2>&1 | tail -1 | cut -c1-330
