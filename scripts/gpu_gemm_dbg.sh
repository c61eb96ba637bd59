for d in 0 1 2; do echo "debug=$d"; I2T_GEMM_DEBUG=$d timeout 300 python scripts/bench_gemm.py 2>&1 | grep -E "B=64|8192\^3|lm_head fwd" | cut -c1-120; done
