mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_decode_mega2.py -m gpu -q --timeout 300 -x -s 2>&1 | tail -30 > gpurun_out/t_mega2.log
cat gpurun_out/t_mega2.log | tail -25
timeout 300 python scripts/trace_mega.py bf16 mega2 > gpurun_out/trace_mega2_bf16.log 2>&1; tail -34 gpurun_out/trace_mega2_bf16.log
