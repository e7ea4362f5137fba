python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_step_loop.py -m gpu -q 2>&1 | tail -8
