#!/bin/bash
# round-2l: GPU suite incl. the new D6 two-axis and speculative-contact tests; grid-padding A/B (147 CTAs vs >= 148) for the two solver kernels
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout -k 5 600 python -m pytest tests/test_speculative_contacts.py tests/test_d6_two_angular_axes.py -m gpu -q -x 2>&1 | tail -30 > $O/r2l_new_tests.txt
tail -15 $O/r2l_new_tests.txt
{
for g in 0 148 296; do
echo "=== xpbd NB2_XPBD_MIN_GRID=$g"; NB2_XPBD_MIN_GRID=$g timeout -k 5 120 python scripts/quick_bench.py 4096 8 quad xpbd 2>&1 | tail -2
done
for g in 0 148; do
echo "=== featherstone NB2_FS_MIN_GRID=$g"; NB2_FS_MIN_GRID=$g timeout -k 5 120 python scripts/quick_bench.py 4096 8 quad featherstone 2>&1 | tail -2
done
for v in 1 0 1 0; do
echo "=== xpbd NB2_COLLIDE_LANE_PER_CONTACT=$v"; NB2_COLLIDE_LANE_PER_CONTACT=$v timeout -k 5 120 python scripts/quick_bench.py 4096 8 quad xpbd 2>&1 | tail -2
done
for v in 1 0; do
echo "=== box stacks NB2_COLLIDE_LANE_PER_CONTACT=$v"; NB2_COLLIDE_LANE_PER_CONTACT=$v timeout -k 5 120 python scripts/quick_bench.py 512 8 stacks xpbd 2>&1 | tail -2
done
} > $O/r2l_min_grid_ab.txt 2>&1
cat $O/r2l_min_grid_ab.txt
timeout -k 5 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 > $O/r2l_gpu_tests.txt
tail -5 $O/r2l_gpu_tests.txt
python bench.py --steps 30 --warmup 5 > $O/r2l_bench_n1.json 2> $O/r2l_bench_n1.err; cut -c1-300 $O/r2l_bench_n1.json; tail -3 $O/r2l_bench_n1.err
