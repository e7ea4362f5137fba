python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
P='import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ("value","ms_per_step")}, {k:round(v["ms_per_step"],3) for k,v in d["kernels"].items()})'
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "$P"
