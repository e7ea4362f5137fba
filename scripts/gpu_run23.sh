python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
P='import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ("value","ms_per_step")}, d["step_roofline"]["frac"], {k:round(v["ms_per_step"],3) for k,v in d["kernels"].items()})'
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "$P"
for tx in 16 8; do
touch icar_b200/csrc/mp_simple.cu; make -C icar_b200/csrc EXTRA_NVFLAGS=-DICAR_MP_TX=$tx 2>&1 | grep -E "error" ; echo "TX=$tx"
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "$P"
done
