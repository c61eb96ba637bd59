mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x 2>&1 | tail -3
timeout 300 python scripts/micro_carveout.py 2>&1 | tail -6
timeout 600 python scripts/sweep_decode.py bf16 2>&1 | grep images | tee gpurun_out/sweep_decode_splitk.jsonl
timeout 600 python scripts/profile_decode_seq.py 8 24 2>&1 | tail -25
timeout 600 python scripts/bench_gemm.py 2>&1 | grep shape | cut -c1-200 | tee gpurun_out/bench_gemm_biasstage.jsonl
timeout 600 python scripts/bench_train.py --dtype bf16 --steps 3 --warmup 3 --graph 1 --config gpt2 --batch 32 2>&1 | tail -1 | cut -c1-400
timeout 600 python scripts/bench_train.py --dtype bf16 --steps 3 --warmup 3 --graph 1 --batch 64 2>&1 | tail -1 | cut -c1-400
