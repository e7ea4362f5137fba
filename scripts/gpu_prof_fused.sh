set -x
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
ncu --set full --clock-control none --import-source on -k regex:k_corrector_fused -s 2 -c 1 -o gpurun_out/prof_fused_v2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:k_mp_simple -s 14 -c 1 -o gpurun_out/prof_mp_v2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1
tail -2 gpurun_out/ncu_full2.log | cut -c1-200
ls -la gpurun_out/
