python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s 2>&1 | grep -E "passed|failed|Error|error|fused advect|config 1|assert " | head -30
for cpt in 2 4; do
  ICAR_FUSED_CPT=$cpt timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-230
done
ICAR_FUSED_CPT=2 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 55 -c 22 --csv --log-file gpurun_out/launches_v4.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
python scripts/launch_summary.py gpurun_out/launches_v4.csv
ICAR_FUSED_CPT=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_corrector_fused -s 2 -c 1 -o gpurun_out/prof_fused_v4 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ICAR_FUSED_CPT=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mp_simple -s 14 -c 1 -o gpurun_out/prof_mp_v4 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1
ls gpurun_out | head -30
