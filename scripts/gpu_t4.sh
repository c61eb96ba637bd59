mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x 2>&1 | tail -12
timeout 600 python scripts/bench_train.py --dtype bf16 --steps 1 --warmup 3 --graph 1 --batch 64 --profile 2>&1 | tail -34 | cut -c1-70,130-200 > gpurun_out/prof_train_b64_v2.txt
timeout 600 python scripts/bench_train.py --dtype bf16 --steps 3 --warmup 3 --graph 1 --batch 64 2>&1 | tail -1 | cut -c1-300
