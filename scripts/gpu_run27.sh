python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
P='import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ("value","ms_per_step")}, {k:round(v["ms_per_step"],3) for k,v in d["kernels"].items() if "thompson" in k})'
timeout 300 python bench.py --mp thompson --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "$P"
touch icar_b200/csrc/mp_thompson.cu; make -C icar_b200/csrc ICAR_TH_FMAD=true 2>&1 | grep -E "error"; grep -A3 "k_mp_thompsonILi320" icar_b200/csrc/_obj/mp_thompson.ptxas.log | grep -E "registers|spill"
timeout 300 python bench.py --mp thompson --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "$P"
timeout 900 python -m pytest tests/test_gpu_thompson.py -m gpu -q 2>&1 | tail -5
