mkdir -p gpurun_out/v2
timeout 600 python -m pytest tests/test_gpu_step_loop.py -m gpu -q -x 2>&1 | tail -5
B="--mp thompson --steps 10 --warmup 3 --no-cpu-baseline --no-parity"
for v in occ3 tx16; do
  timeout 300 python scripts/dev/bench_variant.py icar_b200/_variants/libicar_b200_$v.so $B > gpurun_out/v2/th_$v.json 2> gpurun_out/v2/th_$v.err
done
timeout 300 python scripts/dev/bench_variant.py icar_b200/_variants/libicar_b200_occ3.so $B --nx 512 --ny 512 --nz 60 > gpurun_out/v2/th60_occ3.json 2> gpurun_out/v2/th60_occ3.err
for f in gpurun_out/v2/*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["ms_per_step"], {k:round(v["ms_per_step"],3) for k,v in d.get("kernels").items()})
except Exception as e:
    print("ERR", e)
PY
done
timeout 600 python scripts/dev/bench_variant.py icar_b200/_variants/libicar_b200_tx16.so pytest tests/test_gpu_thompson.py -m gpu -q -x 2>&1 | tail -3
