cd scripts/probe && nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tma_probe tma_probe.cu -lcuda 2>&1 | tail -3
for v in 0 1 2 3; do ./tma_probe $v 0 1; done
./tma_probe 3 -1 1; ./tma_probe 3 40 1; ./tma_probe 3 0 -1; ./tma_probe 3 0 9
