timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q --timeout 600 -x -k "sampler or nucleus or gpt2hf or topk" 2>&1 | tail -15
