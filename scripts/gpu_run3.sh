set -x
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused or decomposition or golden_hill or config1" 2>&1 | tail -15
for cpt in 4 2; do
  ICAR_FUSED_CPT=$cpt timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-330
done
ICAR_FUSED_CPT=2 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 24 --csv --log-file gpurun_out/launches_v3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
python scripts/launch_summary.py gpurun_out/launches_v3.csv
