#!/bin/bash
# round-2d GPU pass: full GPU suite (all failures listed), kernel timings of the three fused kernels, contact-cache A/B, ncu of featherstone
cd "$(dirname "$0")/.."
O=gpurun_out
timeout -k 5 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 > $O/r2d_gpu_tests.txt
cat $O/r2d_gpu_tests.txt
{
echo "=== xpbd default"; timeout -k 5 120 python scripts/quick_bench.py 4096 8 quad xpbd 2>&1 | tail -2
echo "=== xpbd no contact cache"; NB2_XPBD_CONTACT_CACHE=0 timeout -k 5 120 python scripts/quick_bench.py 4096 8 quad xpbd 2>&1 | tail -2
echo "=== collide warps=1"; NB2_COLLIDE_WARPS=1 timeout -k 5 120 python scripts/quick_bench.py 4096 8 quad xpbd 2>&1 | tail -2
echo "=== featherstone default (W auto, sync 1)"; timeout -k 5 120 python scripts/quick_bench.py 4096 8 quad featherstone 2>&1 | tail -2
echo "=== featherstone W=14 sync 0"; NB2_FS_PHASE_SYNC=0 timeout -k 5 120 python scripts/quick_bench.py 4096 8 quad featherstone 2>&1 | tail -2
echo "=== featherstone W=1"; NB2_FS_WARPS=1 timeout -k 5 120 python scripts/quick_bench.py 4096 8 quad featherstone 2>&1 | tail -2
echo "=== featherstone W=4"; NB2_FS_WARPS=4 timeout -k 5 120 python scripts/quick_bench.py 4096 8 quad featherstone 2>&1 | tail -2
echo "=== box stacks"; timeout -k 5 120 python scripts/quick_bench.py 512 8 stacks xpbd 2>&1 | tail -2
} > $O/r2d_kernels.txt 2>&1
cat $O/r2d_kernels.txt
timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:featherstone_step_kernel -s 250 -c 1 -f -o $O/r2d_featherstone python scripts/quick_bench.py 4096 8 quad featherstone > $O/r2d_ncu.log 2>&1
tail -2 $O/r2d_ncu.log
