#!/bin/bash
# round-2f, 2 GPUs: the 2-GPU parity test (peer gather == NCCL == monolithic oracle) and bench lines at N=2 (peer vs NCCL gather)
cd "$(dirname "$0")/.."
O=gpurun_out
nvidia-smi topo -m > $O/r2f_topo.txt 2>&1
timeout -k 5 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_featherstone_parity.py tests/test_gpu_featherstone_tile.py -m gpu -q 2>&1 | tail -30 > $O/r2f_tests.txt
cat $O/r2f_tests.txt
for g in peer nccl; do
  timeout -k 5 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu-baseline --gather $g > $O/r2f_bench_n2_$g.json 2> $O/r2f_bench_n2_$g.err
  cut -c1-200 $O/r2f_bench_n2_$g.json; tail -3 $O/r2f_bench_n2_$g.err
done
NB2_PEER_MEMOPS=1 timeout -k 5 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu-baseline --gather peer > $O/r2f_bench_n2_peer_memops.json 2> $O/r2f_bench_n2_peer_memops.err
cut -c1-200 $O/r2f_bench_n2_peer_memops.json; tail -3 $O/r2f_bench_n2_peer_memops.err
python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $O/r2f_bench_n1.json 2>/dev/null; cut -c1-200 $O/r2f_bench_n1.json
