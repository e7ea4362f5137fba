mkdir -p gpurun_out/v4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
(time timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8) 2>&1 | tee gpurun_out/v4/pytest_gpu.log
B="--mp thompson --steps 10 --warmup 3 --no-cpu-baseline --no-parity"
for nz in 40 50 60; do
  timeout 300 python bench.py $B --nz $nz > gpurun_out/v4/th${nz}.json 2> gpurun_out/v4/th${nz}.err
done
for f in gpurun_out/v4/*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["ms_per_step"], {k:round(v["ms_per_step"],3) for k,v in d.get("kernels").items()})
except Exception as e:
    print("ERR", e)
PY
done
