# ncu evidence for profiles/: launch lists (gpu__time_duration) and one --set full capture of the top kernels.
#   /usr/local/graft/bin/gpurun --timeout 1800 -- 'bash scripts/gpu_profile.sh'
#   then: python scripts/ncu_summary.py gpurun_out/top_simple.ncu-rep 10485760 > profiles/<round>_top_kernels.txt
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 15 --csv --log-file gpurun_out/launches_simple.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
python scripts/launch_summary.py gpurun_out/launches_simple.csv
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_corrector_fused|k_mp_simple|k_upwind_all' -s 6 -c 3 -o gpurun_out/top_simple python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 15 --csv --log-file gpurun_out/launches_thompson.csv python bench.py --mp thompson --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
python scripts/launch_summary.py gpurun_out/launches_thompson.csv
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_mp_thompson' -s 3 -c 1 -o gpurun_out/top_thompson python bench.py --mp thompson --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log
