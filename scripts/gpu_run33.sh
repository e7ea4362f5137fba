python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
P='import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ("value","ms_per_step")}, {k:round(v["ms_per_step"],3) for k,v in d["kernels"].items() if "thompson" in k})'
for tx in 16 4; do
touch icar_b200/csrc/mp_thompson.cu; make -C icar_b200/csrc EXTRA_NVFLAGS=-DICAR_TH_TX=$tx 2>&1 | grep -E "error"; echo "TX=$tx"; grep -E "registers" icar_b200/csrc/_obj/mp_thompson.ptxas.log | tail -3
timeout 300 python bench.py --mp thompson --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "$P"
done
