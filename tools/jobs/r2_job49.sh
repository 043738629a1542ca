#!/bin/bash
# round 2, GPU job 49 (2 GPUs): the bench line under torchrun at N=2 on the final tree; reference arm under torchrun
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 ) > gpurun_out/j49_bench_2gpu.json 2> gpurun_out/j49_bench_2gpu.err
