mkdir -p gpurun_out/v5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
(time timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8) 2>&1 | tee gpurun_out/v5/pytest_gpu.log
timeout 900 python -m pytest tests/test_gpu_thompson.py -m gpu -q -s -k reference_source 2>&1 | grep "columns set aside" | tee -a gpurun_out/v5/pytest_gpu.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/v5/bench_simple.json 2> gpurun_out/v5/bench_simple.err
timeout 600 python bench.py --mp thompson --steps 20 --warmup 5 > gpurun_out/v5/bench_thompson.json 2> gpurun_out/v5/bench_thompson.err
B="python bench.py --no-cpu-baseline --no-parity --e2e-bands 1"
TAG=r2b
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 15 --csv --log-file gpurun_out/${TAG}_launches_thompson.csv $B --mp thompson --steps 2 --warmup 3 > gpurun_out/ncu_b.log 2>&1
python scripts/launch_summary.py gpurun_out/${TAG}_launches_thompson.csv
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_mp_thompson' -s 3 -c 1 -o gpurun_out/${TAG}_top_thompson $B --mp thompson --steps 1 --warmup 1 > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log
for f in gpurun_out/v5/bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["step_roofline"]["frac"], (d.get("parity") or {}).get("bitwise_equal"), {k:round(v["ms_per_step"],3) for k,v in d.get("kernels").items()})
except Exception as e:
    print("ERR", e)
PY
done
