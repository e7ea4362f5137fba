mkdir -p gpurun_out/v3
timeout 900 python -m pytest tests/test_gpu_thompson.py -m gpu -q -x 2>&1 | tail -4
B="--mp thompson --steps 10 --warmup 3 --no-cpu-baseline --no-parity"
timeout 300 python bench.py $B > gpurun_out/v3/th40_default.json 2> gpurun_out/v3/th40_default.err
for nz in 50 60; do
  timeout 300 python bench.py $B --nz $nz > gpurun_out/v3/th${nz}_default.json 2> gpurun_out/v3/th${nz}_default.err
  for v in hi2 tx16hi; do
    timeout 300 python scripts/dev/bench_variant.py icar_b200/_variants/libicar_b200_$v.so $B --nz $nz > gpurun_out/v3/th${nz}_$v.json 2> gpurun_out/v3/th${nz}_$v.err
  done
done
for f in gpurun_out/v3/*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["ms_per_step"], {k:round(v["ms_per_step"],3) for k,v in d.get("kernels").items()})
except Exception as e:
    print("ERR", e)
PY
done
