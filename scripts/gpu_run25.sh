python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
P='import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ("value","ms_per_step")}, {k:round(v["ms_per_step"],3) for k,v in d["kernels"].items() if "upwind" in k})'
for mb in 4 6 8; do
touch icar_b200/csrc/mpdata_staged.cu; make -C icar_b200/csrc EXTRA_NVFLAGS=-DICAR_MARCH_MINB=$mb 2>&1 | grep -E "error"; grep -A3 "k_upwind_marchILi5" icar_b200/csrc/_obj/mpdata_staged.ptxas.log | grep -E "registers|spill"
for kc in 40 20 10 5; do
echo "MINB=$mb kchunk=$kc"; ICAR_MARCH_KCHUNK=$kc timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "$P"
done; done
