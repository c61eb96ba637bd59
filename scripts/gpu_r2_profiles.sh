# round 2 profiles: launch list of the bench command, full capture of the decode kernel (traffic + stall mix)
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train --no-eager-ref > gpurun_out/r02_bench_short.json 2> gpurun_out/r02_bench_short.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train --no-eager-ref > gpurun_out/r02_ncu_launches.log 2>&1
echo "launch list rc=$?"
python scripts/profile_mega3.py 64 > gpurun_out/r02_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:decode_mega3_kernel -s 1 -c 1 -o gpurun_out/r02_mega3_full python scripts/profile_mega3.py 64 > gpurun_out/r02_ncu_full.log 2>&1
echo "full capture rc=$?"; tail -2 gpurun_out/r02_ncu_full.log
