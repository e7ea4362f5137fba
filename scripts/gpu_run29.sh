python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["n_gpus"], d["config"]["decomposition"], {k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["ms_per_step"], d["step_roofline"]["frac"], {k:round(v["ms_per_step"],3) for k,v in d["kernels"].items()})'
for n in 8 4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 100 --warmup 10 2>&1 | grep '"metric"' | tail -1 | python -c "$P"
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --slabs x --steps 100 --warmup 10 2>&1 | grep '"metric"' | tail -1 | python -c "$P"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 8 --mp thompson --steps 20 --warmup 5 2>&1 | grep '"metric"' | tail -1 | python -c "$P"
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "4-y" 2>&1 | tail -3
