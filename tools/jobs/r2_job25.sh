#!/bin/bash
# round 2, GPU job 25 (8 GPUs): the bench line under torchrun at N=8 = the 4096-member C4 target, and the reference arm
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 ) > gpurun_out/j25_bench_8gpu.json 2> gpurun_out/j25_bench_8gpu.err
nvidia-smi --query-gpu=index,name,memory.used --format=csv > gpurun_out/j25_smi.txt 2>&1
