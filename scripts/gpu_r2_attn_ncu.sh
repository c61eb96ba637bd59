# round 2: evidence for the tcgen05 attention kernels -- timings (plain run first), then one full ncu capture per kernel
mkdir -p gpurun_out
python scripts/bench_attn.py --batches 8,64 > gpurun_out/r02_bench_attn.jsonl 2> gpurun_out/r02_bench_attn.err
echo "bench_attn rc=$?"; cat gpurun_out/r02_bench_attn.jsonl
python scripts/bench_attn.py --batches 64 --iters 3 > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attn_bwd_tc5r_kernel -s 3 -c 1 -o gpurun_out/r02_attn_bwd_tc5r_full python scripts/bench_attn.py --batches 64 --iters 3 > gpurun_out/r02_ncu_attn_bwd.log 2>&1
echo "bwd capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:attn_fwd_tc5_kernel -s 3 -c 1 -o gpurun_out/r02_attn_fwd_tc5_full python scripts/bench_attn.py --batches 64 --iters 3 > gpurun_out/r02_ncu_attn_fwd.log 2>&1
echo "fwd capture rc=$?"
ls -la gpurun_out/*.ncu-rep
