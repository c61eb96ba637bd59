mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x 2>&1 | tail -3
for b in 8 64; do
  timeout 600 python scripts/bench_train.py --dtype bf16 --steps 3 --warmup 3 --graph 1 --batch $b 2>&1 | tail -1 | cut -c1-400 | tee -a gpurun_out/train_dropout.jsonl
  timeout 600 python scripts/bench_train.py --dtype bf16 --steps 3 --warmup 3 --graph 1 --batch $b --no-dropout 2>&1 | tail -1 | cut -c1-400 | tee -a gpurun_out/train_dropout.jsonl
done
timeout 600 python scripts/bench_train.py --dtype bf16 --steps 3 --warmup 3 --graph 1 --config gpt2 --batch 32 2>&1 | tail -1 | cut -c1-400 | tee -a gpurun_out/train_dropout.jsonl
timeout 600 python scripts/bench_train.py --dtype bf16 --steps 3 --warmup 3 --graph 1 --config gpt2 --batch 32 --no-dropout 2>&1 | tail -1 | cut -c1-400 | tee -a gpurun_out/train_dropout.jsonl
timeout 600 python scripts/bench_train.py --dtype bf16 --steps 3 --warmup 3 --graph 1 --moco --batch 64 2>&1 | tail -1 | cut -c1-400 | tee -a gpurun_out/train_dropout.jsonl
