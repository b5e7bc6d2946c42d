#!/bin/bash
# bench.py under torchrun on N GPUs of one box:  bash profiles/r2_multi.sh N
N=$1
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo_${N}gpu.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu \
  > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
echo rc=$?; cut -c1-400 gpurun_out/r2_bench_${N}gpu.json; tail -3 gpurun_out/r2_bench_${N}gpu.err
