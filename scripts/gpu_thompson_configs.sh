# BASELINE configs[3] (1024x1024x50, Thompson + MPDATA, 8 GPUs; the other slab axis and the decomposition check ride along)
# and configs[4] (2048x2048x60, linear winds once per forcing interval, 8 GPUs, every rank builds only its band of the state).
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 1500 -- 'bash scripts/gpu_thompson_configs.sh'
mkdir -p gpurun_out
P='import sys,json; d=json.loads(sys.stdin.read()); o=d.get("xslab") or d.get("yslab") or {}; print(d["n_gpus"], d["config"]["grid"], {k:d[k] for k in ("value","ms_per_step","gpu_launches")}, "check", (d.get("decomposition_check") or {}).get("bitwise_equal"), "other axis", o.get("slabs"), o.get("ms_per_step"), (o.get("decomposition_check") or {}).get("bitwise_equal"), {k:round(v["ms_per_step"],3) for k,v in d["kernels"].items()}, (d.get("linear_winds") or {}).get("ms"))'
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 700 $T --master-port 29563 bench.py --gpus 8 --mp thompson --nx 1024 --ny 1024 --nz 50 --steps 20 --warmup 5 2> gpurun_out/c3_8.err | grep '"metric"' | tail -1 | tee gpurun_out/c3_8.json | python -c "$P" || tail -5 gpurun_out/c3_8.err
[ -n "$ONLY_C3" ] || timeout 700 $T --master-port 29564 bench.py --gpus 8 --mp thompson --nx 2048 --ny 2048 --nz 60 --steps 10 --warmup 3 --band-state --linear-winds direct --no-other-axis 2> gpurun_out/c4_8.err | grep '"metric"' | tail -1 | tee gpurun_out/c4_8.json | python -c "$P" || tail -5 gpurun_out/c4_8.err
