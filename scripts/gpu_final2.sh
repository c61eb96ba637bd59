mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -15 > gpurun_out/t_all.log; cat gpurun_out/t_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -2 gpurun_out/bench_default.err; cut -c1-1200 gpurun_out/bench_default.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-400 gpurun_out/bench_ref.json
timeout 600 python scripts/sweep_decode.py bf16 2>&1 | grep images | tee gpurun_out/sweep_decode_final.jsonl
timeout 600 python scripts/sweep_decode_hf.py bf16 2>&1 | grep sequences | tee gpurun_out/sweep_decode_hf.jsonl
rm -f gpurun_out/train_final.jsonl
for args in "--batch 8" "--batch 64" "--batch 64 --no-dropout" "--config gpt2 --batch 32" "--moco --batch 64"; do
  timeout 600 python scripts/bench_train.py --dtype bf16 --steps 3 --warmup 3 --graph 1 $args 2>&1 | tail -1 | cut -c1-400 | tee -a gpurun_out/train_final.jsonl
done
timeout 600 python scripts/bench_train.py --dtype bf16 --steps 1 --warmup 3 --graph 1 --batch 64 --profile 2>&1 | tail -34 | cut -c1-70,130-200 > gpurun_out/prof_train_b64_v2.txt
timeout 600 python scripts/profile_decode_seq.py 8 30 2>&1 | tail -37 > gpurun_out/decode_timeline_b64.txt
