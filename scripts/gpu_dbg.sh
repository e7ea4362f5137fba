set -x
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v "^$" | head -60
