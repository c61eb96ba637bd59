mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_training.py -m gpu -q --timeout 240 2>&1 | tail -150 > gpurun_out/t_all.log
grep -E "passed|failed|^FAILED|^E  " gpurun_out/t_all.log | head -40
for dt in ${DTYPES:-fp32 bf16}; do
  timeout 300 python bench.py --steps 5 --warmup 3 --dtype $dt --no-cpu-baseline > gpurun_out/bench_${dt}.json 2> gpurun_out/bench_${dt}.err
  python - <<PY || tail -5 gpurun_out/bench_${dt}.err
import json
d=json.load(open("gpurun_out/bench_${dt}.json"))
print("$dt", d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["us_per_launch"], d["roofline"]["frac"], d["roofline"]["dominant_kernel"])
PY
done
