# launch list + one --set full capture of the default step's kernels (the first half of scripts/gpu_profile.sh)
TAG=${1:-r2c}
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-parity --e2e-bands 1"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 15 --csv --log-file gpurun_out/${TAG}_launches_simple.csv $B --steps 2 --warmup 3 > gpurun_out/ncu_b.log 2>&1
python scripts/launch_summary.py gpurun_out/${TAG}_launches_simple.csv
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_corrector_fused|k_mp_simple|k_upwind_all|k_face_coefficients|k_wind_normalise' -s 10 -c 5 -o gpurun_out/${TAG}_top_simple $B --steps 1 --warmup 1 > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log
