for b in 8 64; do timeout 600 python scripts/bench_train.py --dtype bf16 --steps 3 --warmup 2 --graph 1 --batch $b 2>&1 | tail -1 | cut -c1-400; done
timeout 600 python scripts/bench_train.py --dtype bf16 --steps 3 --warmup 2 --graph 1 --config gpt2 --batch 32 2>&1 | tail -1 | cut -c1-400
timeout 600 python scripts/bench_train.py --dtype bf16 --steps 1 --warmup 2 --graph 0 --batch 64 --profile > gpurun_out/train_profile_bf16_b64.log 2>&1; head -34 gpurun_out/train_profile_bf16_b64.log | cut -c1-90,150-230
