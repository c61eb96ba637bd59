N=8
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 5 --warmup 3 2>/dev/null | tail -1 | cut -c1-330
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 scripts/bench_train.py --dtype bf16 --steps 3 --warmup 3 --graph 1 --batch 64 2>&1 | tail -1 | cut -c1-300
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 scripts/bench_train.py --dtype bf16 --steps 3 --warmup 3 --graph 1 --config gpt2 --batch 32 2>&1 | tail -1 | cut -c1-300
