python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 600 python bench.py --mp thompson --steps 20 --warmup 5 > gpurun_out/bench_thompson_v1.json 2> gpurun_out/bench_thompson_v1.err; tail -1 gpurun_out/bench_thompson_v1.json; tail -3 gpurun_out/bench_thompson_v1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 30 --csv --log-file gpurun_out/launches_thompson_v1.csv python bench.py --mp thompson --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
python scripts/launch_summary.py gpurun_out/launches_thompson_v1.csv
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_mp_thompson' -s 3 -c 1 -o gpurun_out/r1_thompson_v1 python bench.py --mp thompson --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
