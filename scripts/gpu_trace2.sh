mkdir -p gpurun_out
timeout 300 python scripts/trace_mega.py bf16 mega2 > gpurun_out/trace_mega2_bf16.log 2>&1; tail -34 gpurun_out/trace_mega2_bf16.log
