python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -3
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then
    timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['ms_per_step'], d['roofline']['frac'])"
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 50 --warmup 5 2>&1 | grep '"metric"' | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['ms_per_step'], d['roofline']['frac'])"
  fi
done
