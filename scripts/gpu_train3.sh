for b in 32 64; do timeout 600 python scripts/bench_train.py --dtype bf16 --steps 3 --warmup 2 --graph 1 --batch $b 2>&1 | tail -1 | cut -c1-400; done
timeout 600 python scripts/bench_train.py --dtype bf16 --steps 3 --warmup 2 --graph 1 --config gpt2 --batch 32 2>&1 | tail -1 | cut -c1-400
timeout 600 python scripts/bench_train.py --dtype bf16 --steps 3 --warmup 2 --graph 1 --moco 2>&1 | tail -1 | cut -c1-400
