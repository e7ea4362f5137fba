python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
P='import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["ms_per_step"], d["roofline"]["frac"])'
for c in "" 10 25 50; do
echo "carveout=$c"; env ${c:+ICAR_TH_CARVEOUT=$c} timeout 300 python bench.py --mp thompson --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "$P"
done
