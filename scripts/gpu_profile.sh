# ncu evidence for profiles/: launch lists (gpu__time_duration) and one --set full capture of the top kernels.
#   /usr/local/graft/bin/gpurun --timeout 1800 -- 'bash scripts/gpu_profile.sh r2'
#   then: python scripts/ncu_summary.py gpurun_out/<tag>_top_simple.ncu-rep 10485760 > profiles/<tag>_top_kernels.txt
#         python scripts/ncu_kernels_json.py 10485760 gpurun_out/<tag>_top_simple.ncu-rep gpurun_out/<tag>_top_thompson.ncu-rep > profiles/ncu_kernels.json
TAG=${1:-r2}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
B="python bench.py --no-cpu-baseline --no-parity --e2e-bands 1"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 15 --csv --log-file gpurun_out/${TAG}_launches_simple.csv $B --steps 2 --warmup 3 > gpurun_out/ncu_b.log 2>&1
python scripts/launch_summary.py gpurun_out/${TAG}_launches_simple.csv
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_corrector_fused|k_mp_simple|k_upwind_all|k_face_coefficients|k_wind_normalise' -s 10 -c 5 -o gpurun_out/${TAG}_top_simple $B --steps 1 --warmup 1 > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 15 --csv --log-file gpurun_out/${TAG}_launches_thompson.csv $B --mp thompson --steps 2 --warmup 3 > gpurun_out/ncu_b.log 2>&1
python scripts/launch_summary.py gpurun_out/${TAG}_launches_thompson.csv
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_mp_thompson' -s 3 -c 1 -o gpurun_out/${TAG}_top_thompson $B --mp thompson --steps 1 --warmup 1 > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log
