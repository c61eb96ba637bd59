mkdir -p gpurun_out
python scripts/profile_kernels.py gemm && timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_pair -s 1 -c 1 -f -o gpurun_out/prof_gemm_pair_k768 python scripts/profile_kernels.py gemm > gpurun_out/ncu_gemm1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_pair -s 4 -c 1 -f -o gpurun_out/prof_gemm_pair_8192 python scripts/profile_kernels.py gemm > gpurun_out/ncu_gemm2.log 2>&1
python scripts/profile_kernels.py attn && timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_tc5 -s 1 -c 1 -f -o gpurun_out/prof_attn_fwd_tc5 python scripts/profile_kernels.py attn > gpurun_out/ncu_attn1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_tc -s 1 -c 1 -f -o gpurun_out/prof_attn_bwd_tc python scripts/profile_kernels.py attn > gpurun_out/ncu_attn2.log 2>&1
tail -2 gpurun_out/ncu_gemm1.log gpurun_out/ncu_gemm2.log gpurun_out/ncu_attn1.log gpurun_out/ncu_attn2.log | cut -c1-160
