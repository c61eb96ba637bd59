mkdir -p gpurun_out
python scripts/profile_decode.py bf16 3 > gpurun_out/full_plain_bf16.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:decode_mega -s 1 -c 1 -o gpurun_out/prof_mega_bf16 python scripts/profile_decode.py bf16 3 > gpurun_out/ncu_full_mega.log 2>&1
tail -3 gpurun_out/ncu_full_mega.log
I2T_DECODE=kernels python scripts/profile_decode.py fp32 3 > gpurun_out/full_plain_fp32.log 2>&1 &&
I2T_DECODE=kernels ncu --set full --clock-control none --import-source on -k regex:dec_linear -s 121 -c 1 -o gpurun_out/prof_lmhead_fp32 python scripts/profile_decode.py fp32 3 > gpurun_out/ncu_full_lm.log 2>&1
tail -3 gpurun_out/ncu_full_lm.log
ls -la gpurun_out/*.ncu-rep
