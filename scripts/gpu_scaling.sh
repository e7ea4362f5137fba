# N GPUs of one box: the strong-scaling bench lines (one process per GPU; the halo exchange stores into peer-mapped
# memory over NVLink) with their decomposition checks, and the multi-process decomposition-invariance tests.
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 900 -- 'bash scripts/gpu_scaling.sh 8 4 2'
#   (SKIP_SIMPLE=1 / SKIP_THOMPSON=1 / SKIP_TESTS=1 shorten it: an N-GPU box is charged N x its time; GRAPH=1 replays the steps from
#    a CUDA graph; EXTRA="--nx 1024 --ny 1024 --nz 50" times another grid)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
P='import sys,json; d=json.loads(sys.stdin.read()); o=d.get("xslab") or d.get("yslab") or {}; print(d["n_gpus"], {k:d[k] for k in ("value","ms_per_step","gpu_launches")}, "check", (d.get("decomposition_check") or {}).get("bitwise_equal"), "other axis", o.get("slabs"), o.get("ms_per_step"), (o.get("decomposition_check") or {}).get("bitwise_equal"), "e2e ms", round(d["e2e"]["ms_per_step"],2), {k:round(v["ms_per_step"],3) for k,v in d["kernels"].items()})'
for n in "$@"; do
[ -n "$SKIP_SIMPLE" ] || timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 100 --warmup 10 ${GRAPH:+--graph} $EXTRA 2>gpurun_out/scale_simple_$n.err | grep '"metric"' | tail -1 | tee gpurun_out/scale_simple_$n${TAG}.json | python -c "$P" || tail -5 gpurun_out/scale_simple_$n.err
[ -n "$SKIP_THOMPSON" ] || timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n bench.py --gpus $n --mp thompson --steps 20 --warmup 5 ${GRAPH:+--graph} $EXTRA 2>gpurun_out/scale_thompson_$n.err | grep '"metric"' | tail -1 | tee gpurun_out/scale_thompson_$n${TAG}.json | python -c "$P" || tail -5 gpurun_out/scale_thompson_$n.err
done
[ -n "$SKIP_TESTS" ] || timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -3
