mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bf16_mega2.json 2> gpurun_out/bench_bf16_mega2.err; tail -3 gpurun_out/bench_bf16_mega2.err; cat gpurun_out/bench_bf16_mega2.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_mega2 -s 1 -c 1 -f -o gpurun_out/prof_mega2_bf16 python scripts/trace_mega.py bf16 mega2 > gpurun_out/ncu_full_mega2.log 2>&1; tail -3 gpurun_out/ncu_full_mega2.log
