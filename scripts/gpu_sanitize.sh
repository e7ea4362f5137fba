# One GPU box: compute-sanitizer over the small-grid paths ("race detection / sanitizers").
# memcheck on smoke(), the step-loop kernels and the Thompson instantiations, racecheck + initcheck on smoke()
# (shared-memory exchange rows / TMA stages of the fused corrector, the column reductions of the mp kernels).
# Run it BEFORE timing a new kernel: an out-of-bounds store that happens to land in padding passes the parity
# tests and corrupts a neighbouring field at another size.  ~5 minutes of box time.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash scripts/gpu_sanitize.sh'
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
run() {  # name, tool, command...
    name=$1; tool=$2; shift 2
    timeout 600 $CS --tool $tool --error-exitcode 9 --print-limit 20 "$@" > gpurun_out/sanitize_$name.log 2>&1
    echo "$name ($tool): exit $? ; $(grep -c 'ERROR SUMMARY' gpurun_out/sanitize_$name.log) summaries ; $(grep 'ERROR SUMMARY' gpurun_out/sanitize_$name.log | tail -1)"
}
run smoke_mem   memcheck  python -c "import __graft_entry__ as g; g.smoke()"
run smoke_race  racecheck python -c "import __graft_entry__ as g; g.smoke()"
run smoke_init  initcheck python -c "import __graft_entry__ as g; g.smoke()"
run step_loop   memcheck  python -m pytest tests/test_gpu_step_loop.py -m gpu -q -x
run thompson    memcheck  python -m pytest tests/test_gpu_thompson.py -m gpu -q -x -k "reference_source or 100_levels or tile_and_optional or clipped"
run thompson_race racecheck python -m pytest tests/test_gpu_thompson.py -m gpu -q -x -k "reference_source"
[ -n "$SHORT" ] && { for f in gpurun_out/sanitize_*.log; do tail -n 3 $f; done | tail -n 40; exit 0; }
run linear_winds memcheck python -m pytest tests/test_gpu_linear_winds.py -m gpu -q -x -k "not full_size"
run restart      memcheck python -m pytest tests/test_gpu_restart.py -m gpu -q -x
tail -5 gpurun_out/sanitize_*.log | tail -60
