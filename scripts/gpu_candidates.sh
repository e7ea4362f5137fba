# One GPU box: run everything that was built but has not run on hardware yet ("Environment hooks"),
# then time each candidate against the default.  ~2 minutes of box time.
#   /usr/local/graft/bin/gpurun --timeout 400 -- 'bash scripts/gpu_candidates.sh'
# Two GPUs (multi-image graph replay with the NCCL exchange inside the captured pair):
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 300 -- 'ICAR_TEST_GRAPH_MULTI=1 python -m pytest tests/test_gpu_multi.py -m gpu -q'
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
ICAR_TEST_UNVALIDATED=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q \
    -k "candidate or 52_level or packed" > gpurun_out/candidates_tests.log 2>&1
ICAR_TEST_UNVALIDATED=1 timeout 120 python -m pytest tests/test_gpu_step_loop.py -m gpu -q -k "step_host or host_pipeline" >> gpurun_out/candidates_tests.log 2>&1
ICAR_TEST_UNVALIDATED=1 timeout 300 python -m pytest tests/test_cpp_host.py tests/test_gpu_step_graph.py -m gpu -q >> gpurun_out/candidates_tests.log 2>&1
ICAR_TEST_UNVALIDATED=1 timeout 400 python -m pytest tests/test_gpu_linear_winds.py tests/test_gpu_restart.py -m gpu -q >> gpurun_out/candidates_tests.log 2>&1
ICAR_TEST_UNVALIDATED=1 timeout 400 python -m pytest tests/test_gpu_thompson.py -m gpu -q -k math2 >> gpurun_out/candidates_tests.log 2>&1
tail -15 gpurun_out/candidates_tests.log
P='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], "ms/step", round(d["ms_per_step"],4), "e2e ms", round(d["e2e"]["ms_per_step"],2), {k:round(v["ms_per_step"],4) for k,v in d["kernels"].items()})'
run() {  # name, env assignments..., then -- bench arguments
    name=$1; shift
    envs=()
    while [ "$1" != "--" ]; do envs+=("$1"); shift; done
    shift
    env "${envs[@]}" timeout 120 python bench.py --steps 100 --warmup 5 --no-cpu-baseline "$@" \
        > gpurun_out/cand_$name.json 2> gpurun_out/cand_$name.err && python -c "$P" $name < gpurun_out/cand_$name.json \
        || { echo "$name FAILED"; tail -3 gpurun_out/cand_$name.err; }
}
run default -- 
run pass1_packed ICAR_UPWIND_PACKED=1 --
run corr_scalar ICAR_CORR_SCALAR=1 --
run graph -- --graph
run derive_diagnostics -- --derive-diagnostics     # look at e2e in gpurun_out/cand_*.json for these four
run e2e_bands4 -- --e2e-bands 4
run e2e_bands8 -- --e2e-bands 8
run e2e_bands8_derive -- --e2e-bands 8 --derive-diagnostics
run linear_winds -- --linear-winds                 # look at linear_winds.ms in gpurun_out/cand_linear_winds.json
run linear_winds_spatial -- --linear-winds spatial
run thompson_default -- --mp thompson --steps 20
run thompson_math2 ICAR_TH_MATH2=1 -- --mp thompson --steps 20
run nz50_default -- --nz 50
run nz50_two_stages ICAR_FUSED_NZP52=1 -- --nz 50
run nz50_packed ICAR_CORR_V2=1 -- --nz 50
run nz50_packed_two_stages ICAR_CORR_V2=1 ICAR_FUSED_NZP52=1 -- --nz 50
run nz60_default -- --nz 60
run nz60_packed ICAR_CORR_V2=1 -- --nz 60
run nz60_packed_two_stages ICAR_CORR_V2=1 ICAR_FUSED_NZP60=1 -- --nz 60
