#!/bin/bash
# round 2, GPU job 12 (2 GPUs): the bench line under torchrun at N=2 (C4 ensemble, rain shared over NCCL), reference arm
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 ) > gpurun_out/j12_bench_2gpu.json 2> gpurun_out/j12_bench_2gpu.err
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --workload c2 ) > gpurun_out/j12_bench_c2_2gpu.json 2> gpurun_out/j12_bench_c2_2gpu.err
ls -la gpurun_out > gpurun_out/j12_ls.txt
