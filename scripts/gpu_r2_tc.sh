# tcgen05 linear stages of the dataflow decode kernel: parity tests (guarded by a timeout), bench A/B against the mma.sync path
mkdir -p gpurun_out
T=${TAG:-r2m}
I2T_M3_TC=1 timeout 600 python -m pytest tests/test_gpu_decode_mega3.py -m gpu -x -q -s --timeout 300 > gpurun_out/${T}_t_mega3_tc.log 2>&1
echo "mega3 tc tests rc=$?"; tail -25 gpurun_out/${T}_t_mega3_tc.log
for tc in 1 0; do
  I2T_M3_TC=$tc timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-train --no-eager-ref > gpurun_out/${T}_bench_tc$tc.json 2> gpurun_out/${T}_bench_tc$tc.err
  python - <<PY || tail -5 gpurun_out/${T}_bench_tc$tc.err
import json
d=json.load(open("gpurun_out/${T}_bench_tc$tc.json"))
print("tc $tc", d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["us_per_step"], d["roofline"]["frac"])
PY
done
I2T_M3_TC=1 timeout 300 python scripts/trace_mega3.py 0 > gpurun_out/${T}_trace_tc.txt 2>&1; tail -32 gpurun_out/${T}_trace_tc.txt
