mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q --timeout 300 -x -k "gemm" 2>&1 | tail -5
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x 2>&1 | tail -3
timeout 600 python scripts/sweep_decode.py bf16 2>&1 | grep images | tee gpurun_out/sweep_decode_splitk.jsonl
timeout 300 python scripts/profile_decode_batch.py 8 2>&1 | tail -32 > gpurun_out/prof_decode_b64_splitk.txt
timeout 300 python scripts/profile_decode_batch.py 64 2>&1 | tail -32 > gpurun_out/prof_decode_b512_splitk.txt
timeout 600 python scripts/bench_train.py --dtype bf16 --steps 3 --warmup 3 --graph 1 --config gpt2 --batch 32 2>&1 | tail -1 | cut -c1-400
timeout 600 python scripts/bench_train.py --dtype bf16 --steps 3 --warmup 3 --graph 1 --batch 64 2>&1 | tail -1 | cut -c1-400
