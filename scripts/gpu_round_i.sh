#!/bin/bash
# round-2i (8 GPUs): weak-scaling bench lines with the peer (copy-engine) gather and the NCCL gather, launched exactly like the driver does
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
run() {  # N gather
  timeout -k 5 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) \
    bench.py --gpus $1 --steps 50 --warmup 5 --no-cpu-baseline --gather $2 > $O/r2i_bench_n$1_$2.json 2> $O/r2i_bench_n$1_$2.err
  grep -h '^{' $O/r2i_bench_n$1_$2.json | cut -c1-260
}
nvidia-smi topo -m > $O/r2i_topo.txt 2>&1
run 8 peer
run 8 nccl
run 4 peer
run 2 peer
