mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q --timeout 600 -x -k "large_batch or submodule" 2>&1 | tail -6
timeout 900 python scripts/sweep_decode.py bf16 > gpurun_out/sweep_decode_bf16.jsonl 2>&1; tail -8 gpurun_out/sweep_decode_bf16.jsonl
