#!/bin/bash
# round-2e GPU pass: full GPU suite (new: broad phases, matching, tile GEMM, hull, seeded stacks), tile-GEMM timing + tensor-pipe ncu, bench lines
cd "$(dirname "$0")/.."
O=gpurun_out
timeout -k 5 1500 python -m pytest tests -m gpu -q 2>&1 | tail -60 > $O/r2e_gpu_tests.txt
cat $O/r2e_gpu_tests.txt
{
echo "=== featherstone FP32 path"; timeout -k 5 120 python scripts/quick_bench.py 4096 8 quad featherstone 2>&1 | tail -2
echo "=== featherstone use_tile_gemm (mma.sync TF32x3)"; NB2_TILE=1 timeout -k 5 120 python scripts/quick_bench.py 4096 8 quad featherstone 2>&1 | tail -2
echo "=== xpbd"; timeout -k 5 120 python scripts/quick_bench.py 4096 8 quad xpbd 2>&1 | tail -2
} > $O/r2e_kernels.txt 2>&1
cat $O/r2e_kernels.txt
NB2_TILE=1 timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:featherstone_step_kernel -s 250 -c 1 -f -o $O/r2e_featherstone_tile python scripts/quick_bench.py 4096 8 quad featherstone > $O/r2e_ncu.log 2>&1
tail -2 $O/r2e_ncu.log
for wl in quadruped_xpbd quadruped_featherstone box_stacks_xpbd quadruped_xpbd_stock; do
  python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline > $O/r2e_bench_$wl.json 2> $O/r2e_bench_$wl.err; cut -c1-220 $O/r2e_bench_$wl.json
done
