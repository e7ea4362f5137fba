# One GPU box: smoke, the whole GPU suite (no -x, full log kept), and the headline bench eager vs graph replay.
#   /usr/local/graft/bin/gpurun --timeout 140 -- 'bash scripts/gpu_graph_check.sh'
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 100 python -m pytest tests -m gpu -q > gpurun_out/graph_suite.log 2>&1; tail -6 gpurun_out/graph_suite.log
timeout 60 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_eager.json 2> gpurun_out/bench_eager.err
timeout 60 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --graph > gpurun_out/bench_graph.json 2> gpurun_out/bench_graph.err
python - <<'PY'
import json
for n in ("eager", "graph"):
    try:
        d = json.loads(open(f"gpurun_out/bench_{n}.json").read().strip().splitlines()[-1])
        print(n, "ms/step", round(d["ms_per_step"], 4), "launches", d["gpu_launches"], "e2e ms", round(d["e2e"]["ms_per_step"], 2), d["clocks"])
    except Exception as e:
        print(n, "failed:", e, open(f"gpurun_out/bench_{n}.err").read()[-600:])
PY
