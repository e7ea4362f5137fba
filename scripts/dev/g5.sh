mkdir -p gpurun_out/v5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
(timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6) 2>&1 | tee gpurun_out/v5/pytest_gpu.log
timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/v5/simple_new.json 2> gpurun_out/v5/simple_new.err
timeout 300 python scripts/dev/bench_variant.py icar_b200/_variants/libicar_b200_head.so --steps 50 --warmup 5 > gpurun_out/v5/simple_head.json 2> gpurun_out/v5/simple_head.err
timeout 300 python bench.py --mp thompson --steps 20 --warmup 5 > gpurun_out/v5/th_new.json 2> gpurun_out/v5/th_new.err
timeout 300 python scripts/dev/bench_variant.py icar_b200/_variants/libicar_b200_head.so --mp thompson --steps 20 --warmup 5 > gpurun_out/v5/th_head.json 2> gpurun_out/v5/th_head.err
for f in gpurun_out/v5/*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["ms_per_step"], {k:round(v["ms_per_step"],3) for k,v in d.get("kernels").items()}, d.get("parity"))
except Exception as e: print("ERR", e)
PY
done
