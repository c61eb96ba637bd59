# ncu on one skinny GEMM (2048 x 768 x 3072, bf16 out): the single-CTA 96-column tile vs the CTA-pair 256 x 128 tile
mkdir -p gpurun_out
python scripts/one_gemm.py 2048 768 3072 nt 5
I2T_GEMM_PAIR_MIN=40 python scripts/one_gemm.py 2048 768 3072 nt 5
ncu --set full --clock-control none -k regex:gemm_tc -s 2 -c 1 -o gpurun_out/r02_gemm_single_2048x768x3072 python scripts/one_gemm.py 2048 768 3072 nt 4 > gpurun_out/r02_ncu_gemm_single.log 2>&1
echo "single rc=$?"
I2T_GEMM_PAIR_MIN=40 ncu --set full --clock-control none -k regex:gemm_tc -s 2 -c 1 -o gpurun_out/r02_gemm_pair_2048x768x3072 python scripts/one_gemm.py 2048 768 3072 nt 4 > gpurun_out/r02_ncu_gemm_pair.log 2>&1
echo "pair rc=$?"
