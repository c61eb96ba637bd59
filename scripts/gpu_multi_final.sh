N=${1:-4}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 5 --warmup 3 2>/dev/null | tail -1 | cut -c1-330
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 scripts/bench_train.py --dtype bf16 --steps 3 --warmup 3 --graph 1 --batch 64 2>&1 | tail -1 | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 scripts/bench_train.py --dtype bf16 --steps 3 --warmup 3 --graph 1 --config gpt2 --batch 32 2>&1 | tail -1 | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29524 scripts/dp_parity.py 2>&1 | tail -1 | cut -c1-200
