python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -3
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["config"]["grid"], {k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["ms_per_step"], d["roofline"]["frac"])'
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "$P"
timeout 300 python bench.py --nx 70 --steps 200 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 | python -c "$P"
timeout 300 python bench.py --nx 134 --steps 200 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 | python -c "$P"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 18 --csv --log-file gpurun_out/launches_nx70.csv python bench.py --nx 70 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
python scripts/launch_summary.py gpurun_out/launches_nx70.csv
