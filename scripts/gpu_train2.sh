mkdir -p gpurun_out
for g in 1 0; do timeout 600 python scripts/bench_train.py --dtype bf16 --steps 4 --warmup 2 --graph $g 2>&1 | tail -2 | cut -c1-400; done
