mkdir -p gpurun_out
timeout 900 python -m pytest ${TESTS:-tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_training.py} -m gpu -q --timeout 300 2>&1 | tail -150 > gpurun_out/t_all.log
grep -E "passed|failed|^FAILED|^E  " gpurun_out/t_all.log | head -40
for mode in ${MODES:-mega kernels}; do for dt in ${DTYPES:-fp32 bf16}; do
  I2T_DECODE=$mode timeout 300 python bench.py --steps 5 --warmup 3 --dtype $dt --no-cpu-baseline > gpurun_out/bench_${dt}_$mode.json 2> gpurun_out/bench_${dt}_$mode.err
  python - <<PY || tail -5 gpurun_out/bench_${dt}_$mode.err
import json
d=json.load(open("gpurun_out/bench_${dt}_$mode.json"))
print("$dt $mode", d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["us_per_launch"], d["roofline"]["frac"], d["gpu_launches"])
PY
done; done
