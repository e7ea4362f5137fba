set -x
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -3
python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" | tail -40
python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_r1_v2.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_v2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
tail -1 gpurun_out/ncu_b.log | cut -c1-300
