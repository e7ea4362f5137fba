python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -3
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches')}, d['roofline']['frac'])"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_upwind_all -s 1 -c 1 -o gpurun_out/prof_upwind_v6 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_face_coef -s 1 -c 1 -o gpurun_out/prof_coef_v6 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full4.log 2>&1
ls gpurun_out/*.ncu-rep
