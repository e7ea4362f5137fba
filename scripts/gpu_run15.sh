python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 900 python scripts/probe/thompson_probe2.py 2>&1 | tail -70
