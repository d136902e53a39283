#!/bin/bash
# round-2c GPU pass: full GPU suite, phase-sync A/B of the default library, ncu capture of xpbd_step_kernel, bench lines
cd "$(dirname "$0")/.."
O=gpurun_out
timeout -k 5 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > $O/r2c_gpu_tests.txt
cat $O/r2c_gpu_tests.txt
{
for W in 4 14; do for S in 0 1 2; do
  echo "=== warps=$W phase_sync=$S"
  NB2_XPBD_WARPS=$W NB2_XPBD_PHASE_SYNC=$S timeout -k 5 120 python scripts/quick_bench.py 4096 8 quad xpbd 2>&1 | tail -2
done; done
echo "=== default, 16384 envs"; timeout -k 5 200 python scripts/quick_bench.py 16384 8 quad xpbd 2>&1 | tail -2
echo "=== default, box stacks 512"; timeout -k 5 200 python scripts/quick_bench.py 512 8 stacks xpbd 2>&1 | tail -2
echo "--- parity phase_sync=2"; NB2_XPBD_PHASE_SYNC=2 timeout -k 5 300 python -m pytest tests/test_gpu_xpbd_parity.py -m gpu -x -q 2>&1 | tail -2
} > $O/r2c_xpbd_ab.txt 2>&1
cat $O/r2c_xpbd_ab.txt
timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:xpbd_step_kernel -s 250 -c 1 -f -o $O/r2c_xpbd python scripts/quick_bench.py 4096 8 quad xpbd > $O/r2c_ncu.log 2>&1
tail -3 $O/r2c_ncu.log
python bench.py --steps 20 --warmup 5 > $O/r2c_bench_n1.json 2> $O/r2c_bench_n1.err; cut -c1-400 $O/r2c_bench_n1.json
python bench.py --impl reference --steps 20 --warmup 5 > $O/r2c_bench_ref.json 2> $O/r2c_bench_ref.err; cut -c1-300 $O/r2c_bench_ref.json
