# re-entry check: smoke, gpu tests, full bench line, launch list, ncu --set full of the two top kernels
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 400 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_r1_v8.json 2> gpurun_out/bench_r1_v8.err; tail -1 gpurun_out/bench_r1_v8.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 24 --csv --log-file gpurun_out/launches_v8.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
python scripts/launch_summary.py gpurun_out/launches_v8.csv
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_corrector_fused|k_mp_simple|k_upwind_all' -s 6 -c 3 -o gpurun_out/r1_v8_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out
