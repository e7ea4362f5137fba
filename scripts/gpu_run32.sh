python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 15 --csv --log-file gpurun_out/launches_v10.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
python scripts/launch_summary.py gpurun_out/launches_v10.csv
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_corrector_fused|k_mp_simple|k_upwind_all' -s 6 -c 3 -o gpurun_out/r1_v10_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 15 --csv --log-file gpurun_out/launches_thompson_v2.csv python bench.py --mp thompson --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
python scripts/launch_summary.py gpurun_out/launches_thompson_v2.csv
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_mp_thompson' -s 3 -c 1 -o gpurun_out/r1_thompson_v2 python bench.py --mp thompson --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log
