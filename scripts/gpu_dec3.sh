mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x 2>&1 | tail -3
timeout 600 python scripts/sweep_decode.py bf16 2>&1 | grep images | tee gpurun_out/sweep_decode_v3.jsonl
timeout 600 python scripts/profile_decode_seq.py 8 20 2>&1 | tail -27
