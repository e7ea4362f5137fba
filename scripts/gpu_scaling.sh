# N GPUs of one box: NCCL decomposition-invariance tests and the strong-scaling bench lines.
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 900 -- 'bash scripts/gpu_scaling.sh 8 4 2'
#   (SKIP_THOMPSON=1 / SKIP_TESTS=1 shorten it: an N-GPU box is charged N x its time;
#    GRAPH=1 steps through icar_physics_steps -- CUDA-graph replay with the NCCL exchange inside the captured pair --
#    after ICAR_TEST_GRAPH_MULTI=1 python -m pytest tests/test_gpu_multi.py -m gpu has passed on 2 GPUs)
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["n_gpus"], d["config"]["decomposition"], {k:d[k] for k in ("value","ms_per_step","gpu_launches")}, "e2e ms", round(d["e2e"]["ms_per_step"],2), "step frac", round(d["step_roofline"]["frac"],4), {k:round(v["ms_per_step"],3) for k,v in d["kernels"].items()})'
for n in "$@"; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 100 --warmup 10 ${GRAPH:+--graph} 2>&1 | grep '"metric"' | tail -1 | tee gpurun_out/scale_simple_$n.json | python -c "$P"
[ -n "$SKIP_THOMPSON" ] || timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n bench.py --gpus $n --mp thompson --steps 20 --warmup 5 ${GRAPH:+--graph} 2>&1 | grep '"metric"' | tail -1 | tee gpurun_out/scale_thompson_$n.json | python -c "$P"
done
[ -n "$SKIP_TESTS" ] || timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -3
