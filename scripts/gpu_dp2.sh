N=${1:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 scripts/dp_parity.py 2>&1 | tail -3 | cut -c1-250
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 scripts/bench_train.py --dtype bf16 --steps 3 --warmup 3 --graph 1 --config gpt2 --batch 32 2>&1 | tail -1 | cut -c1-300
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 scripts/bench_train.py --dtype bf16 --steps 3 --warmup 3 --graph 1 --batch 64 2>&1 | tail -1 | cut -c1-300
