mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x 2>&1 | tail -4 > gpurun_out/t_all.log; cat gpurun_out/t_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -2 gpurun_out/bench_default.err; cat gpurun_out/bench_default.json | cut -c1-3000
