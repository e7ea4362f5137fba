# One GPU box (B200): build, smoke, GPU parity tests, the headline bench line and the Thompson one.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/gpu_check.sh'
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6
timeout 500 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_simple.json 2> gpurun_out/bench_simple.err; tail -1 gpurun_out/bench_simple.json
timeout 500 python bench.py --mp thompson --steps 20 --warmup 5 > gpurun_out/bench_thompson.json 2> gpurun_out/bench_thompson.err; tail -1 gpurun_out/bench_thompson.json
