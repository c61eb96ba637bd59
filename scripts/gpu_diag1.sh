mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x 2>&1 | tail -30 > gpurun_out/t_all.log
grep -E "passed|failed|^FAILED|^E  " gpurun_out/t_all.log | head -20
timeout 300 python scripts/trace_mega.py bf16 > gpurun_out/trace_mega_bf16.log 2>&1; tail -25 gpurun_out/trace_mega_bf16.log
timeout 300 python scripts/bench_gemm.py > gpurun_out/bench_gemm.log 2>&1; cat gpurun_out/bench_gemm.log
for a in 1 0; do I2T_TC_ATTN=$a timeout 600 python scripts/bench_train.py --dtype bf16 --steps 3 --warmup 1 2>&1 | tail -1; done
timeout 600 python scripts/bench_train.py --dtype bf16 --steps 1 --warmup 1 --profile > gpurun_out/train_profile_bf16.log 2>&1; head -40 gpurun_out/train_profile_bf16.log
