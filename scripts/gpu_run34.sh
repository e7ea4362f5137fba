python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["n_gpus"], d["config"]["decomposition"], {k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["ms_per_step"], d["step_roofline"]["frac"], {k:round(v["ms_per_step"],3) for k,v in d["kernels"].items()})'
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 100 --warmup 10 2>&1 | grep '"metric"' | tail -1 | python -c "$P"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --ny 140 --steps 200 --warmup 10 2>&1 | grep '"metric"' | tail -1 | python -c "$P"
