python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1
python scripts/probe/pcie_probe.py
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -3
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-230
for c in 2 5 8; do echo "ICAR_MP_CPT=$c"; ICAR_MP_CPT=$c timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c60-230; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 55 -c 22 --csv --log-file gpurun_out/launches_v5.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
python scripts/launch_summary.py gpurun_out/launches_v5.csv
