python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1
timeout 300 python bench.py --mp thompson --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c 1-2500
