# final state of a round on one B200: smoke, the GPU suite, the two bench lines
mkdir -p gpurun_out/final
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1
(time timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6) 2>&1 | tee gpurun_out/final/pytest_gpu.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/final/bench_simple.json 2> gpurun_out/final/bench_simple.err
timeout 600 python bench.py --mp thompson --steps 20 --warmup 5 > gpurun_out/final/bench_thompson.json 2> gpurun_out/final/bench_thompson.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final/bench_reference.json 2> gpurun_out/final/bench_reference.err
for f in gpurun_out/final/bench_simple.json gpurun_out/final/bench_thompson.json; do echo $f; python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], "e2e", d["e2e"]["ms_per_step"], "roofline", d["roofline"]["frac"], "step", d["step_roofline"]["frac"], {k:round(v["ms_per_step"],3) for k,v in d.get("kernels").items()})
PY
done
python -c "
import json; d=json.loads(open('gpurun_out/final/bench_reference.json').read().strip().splitlines()[-1]); print('reference arm', d['ms_per_step'], d['value'], d['cpu_baseline']['cores'])"
