#!/bin/bash
# 4 GPUs: weak-scaling bench lines with the peer (copy-engine) gather and the NCCL gather, launched like the driver does; 2-GPU test
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
run() {  # N gather
  timeout -k 5 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) \
    bench.py --gpus $1 --steps 50 --warmup 5 --no-cpu-baseline --gather $2 > $O/r2i_bench_n$1_$2.json 2> $O/r2i_bench_n$1_$2.err
  grep -h '^{' $O/r2i_bench_n$1_$2.json | cut -c1-260
  tail -2 $O/r2i_bench_n$1_$2.err
}
nvidia-smi topo -m > $O/r2i_topo.txt 2>&1
run 4 peer
run 4 nccl
run 2 peer
timeout -k 5 300 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -3
