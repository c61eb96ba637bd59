mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_decode_mega2.py -m gpu -q --timeout 300 -x 2>&1 | tail -5
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bf16_mega2.json 2> gpurun_out/bench_bf16_mega2.err; tail -3 gpurun_out/bench_bf16_mega2.err; cat gpurun_out/bench_bf16_mega2.json
I2T_ENCODER_GRAPH=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-400
