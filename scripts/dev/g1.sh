mkdir -p gpurun_out/v1
B="--mp thompson --steps 10 --warmup 3 --no-cpu-baseline --no-parity"
timeout 300 python bench.py $B > gpurun_out/v1/th_default.json 2> gpurun_out/v1/th_default.err
for v in tx4 tx2 tx8p tx4p; do
  timeout 300 python scripts/dev/bench_variant.py icar_b200/_variants/libicar_b200_$v.so $B > gpurun_out/v1/th_$v.json 2> gpurun_out/v1/th_$v.err
done
for v in tx4 tx2; do
  timeout 300 python scripts/dev/bench_variant.py icar_b200/_variants/libicar_b200_$v.so $B --nx 512 --ny 512 --nz 60 > gpurun_out/v1/th60_$v.json 2> gpurun_out/v1/th60_$v.err
done
timeout 300 python bench.py $B --nx 512 --ny 512 --nz 60 > gpurun_out/v1/th60_default.json 2> gpurun_out/v1/th60_default.err
for f in gpurun_out/v1/*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["ms_per_step"], d.get("kernels"))
except Exception as e:
    print("ERR", e)
PY
done
timeout 600 python scripts/dev/bench_variant.py icar_b200/_variants/libicar_b200_tx4.so pytest tests/test_gpu_thompson.py -m gpu -q -x 2>&1 | tail -3
