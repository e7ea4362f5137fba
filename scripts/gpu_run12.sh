python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 24 --csv --log-file gpurun_out/launches_v7.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
python scripts/launch_summary.py gpurun_out/launches_v7.csv
python - <<'PY'
import sys; sys.path.insert(0,'scripts')
import launch_summary as L
seq=L.main('gpurun_out/launches_v7.csv')
for s in seq[:12]: print(s)
PY
