timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x 2>&1 | tail -4
timeout 600 python scripts/bench_train.py --dtype bf16 --steps 3 --warmup 3 --graph 1 --batch 64 2>&1 | tail -1 | cut -c1-300
timeout 600 python scripts/bench_train.py --dtype bf16 --steps 3 --warmup 3 --graph 1 --batch 8 2>&1 | tail -1 | cut -c1-300
timeout 600 python scripts/bench_train.py --dtype bf16 --steps 3 --warmup 3 --graph 1 --config gpt2 --batch 32 2>&1 | tail -1 | cut -c1-300
timeout 600 python scripts/profile_encode_seq.py 2 2>&1 | tail -18 | head -8
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-260
