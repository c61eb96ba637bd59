# round 2: first run of the dataflow decode kernel (mega3): parity tests, bench A/B against mega2, per-stage trace
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_decode_mega3.py -m gpu -x -q -s --timeout 600 > gpurun_out/r2a_t_mega3.log 2>&1
echo "mega3 tests rc=$?"; tail -25 gpurun_out/r2a_t_mega3.log
for mode in mega3 mega2; do
  I2T_DECODE=$mode timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/r2a_bench_$mode.json 2> gpurun_out/r2a_bench_$mode.err
  python - <<PY || tail -5 gpurun_out/r2a_bench_$mode.err
import json
d=json.load(open("gpurun_out/r2a_bench_$mode.json"))
print("$mode", d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["us_per_step"], d["roofline"]["frac"], d["gpu_launches"], d["clocks"])
PY
done
timeout 600 python scripts/trace_mega3.py 0 37 101 147 > gpurun_out/r2a_trace_mega3.txt 2>&1; tail -60 gpurun_out/r2a_trace_mega3.txt
