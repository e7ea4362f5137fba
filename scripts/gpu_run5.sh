python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -s 2>&1 | tail -8
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -4
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-260
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 30 --warmup 5 2>&1 | tail -2 | cut -c1-700
