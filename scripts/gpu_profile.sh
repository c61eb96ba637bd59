mkdir -p gpurun_out
for dt in fp32 bf16; do
python scripts/profile_decode.py $dt 4 > gpurun_out/prof_plain_$dt.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_$dt.csv python scripts/profile_decode.py $dt 4 > gpurun_out/ncu_$dt.log 2>&1
tail -2 gpurun_out/ncu_$dt.log
done
