# round 2: dataflow decode kernel (mega3): parity tests, bench A/B (poll back-off), per-stage trace
mkdir -p gpurun_out
T=${TAG:-r2b}
timeout 900 python -m pytest tests/test_gpu_decode_mega3.py -m gpu -x -q -s --timeout 600 > gpurun_out/${T}_t_mega3.log 2>&1
echo "mega3 tests rc=$?"; tail -25 gpurun_out/${T}_t_mega3.log
for sl in ${SLEEPS:-0 20 100}; do
  I2T_POLL_SLEEP=$sl I2T_DECODE=mega3 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/${T}_bench_mega3_s$sl.json 2> gpurun_out/${T}_bench_mega3_s$sl.err
  python - <<PY || tail -5 gpurun_out/${T}_bench_mega3_s$sl.err
import json
d=json.load(open("gpurun_out/${T}_bench_mega3_s$sl.json"))
print("mega3 sleep $sl", d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["us_per_step"], d["roofline"]["frac"], d["gpu_launches"], d["clocks"])
PY
done
timeout 600 python scripts/trace_mega3.py ${CTAS:-0 101} > gpurun_out/${T}_trace_mega3.txt 2>&1; tail -${TAILN:-70} gpurun_out/${T}_trace_mega3.txt
