#!/bin/bash
# round-2k: parity suite on the unrolled-row / shared 2w library with the two-kernel export default, timings, ncu --set full of
# collide_kernel, contact_export_kernel and xpbd_step_kernel (one launch each, steady state)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout -k 5 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 > $O/r2k_gpu_tests.txt
tail -3 $O/r2k_gpu_tests.txt
{
echo "=== xpbd (default: unrolled rows, split export)"; timeout -k 5 120 python scripts/quick_bench.py 4096 8 quad xpbd 2>&1 | tail -2
echo "=== featherstone"; timeout -k 5 120 python scripts/quick_bench.py 4096 8 quad featherstone 2>&1 | tail -2
echo "=== box stacks"; timeout -k 5 120 python scripts/quick_bench.py 512 8 stacks xpbd 2>&1 | tail -2
} > $O/r2k_kernels.txt 2>&1
cat $O/r2k_kernels.txt
for k in collide_kernel contact_export_kernel xpbd_step_kernel; do
  timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 200 -c 1 -f -o $O/r2k_$k python scripts/quick_bench.py 4096 8 quad xpbd > $O/r2k_ncu_$k.log 2>&1
  tail -2 $O/r2k_ncu_$k.log
done
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/r2k_bench_n1.json 2> $O/r2k_bench_n1.err; cut -c1-200 $O/r2k_bench_n1.json
