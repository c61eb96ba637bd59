mkdir -p gpurun_out
N=${N:-4}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -2 gpurun_out/bench_n$N.err | cut -c1-200; cut -c1-330 gpurun_out/bench_n$N.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 scripts/bench_train.py --dtype bf16 --steps 3 --warmup 2 --graph 1 2>&1 | tail -1 | cut -c1-330
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 scripts/bench_train.py --dtype bf16 --steps 3 --warmup 2 --graph 1 --batch 64 2>&1 | tail -1 | cut -c1-330
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 scripts/bench_train.py --dtype bf16 --steps 3 --warmup 2 --graph 1 --config gpt2 --batch 32 --moco 2>&1 | tail -1 | cut -c1-330
