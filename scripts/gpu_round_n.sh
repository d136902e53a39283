#!/bin/bash
# round-2n: GPU suite after the multi-axis eval_ik / golden additions, large-batch timings, ncu --set full of the Featherstone kernel
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout -k 5 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 > $O/r2n_gpu_tests.txt
tail -4 $O/r2n_gpu_tests.txt
{
for e in 4096 16384 32768; do
echo "=== xpbd envs=$e"; timeout -k 5 200 python scripts/quick_bench.py $e 8 quad xpbd 2>&1 | tail -2
done
echo "=== xpbd fine phase barriers (NB2_XPBD_PHASE_SYNC=2)"; NB2_XPBD_PHASE_SYNC=2 timeout -k 5 200 python scripts/quick_bench.py 4096 8 quad xpbd 2>&1 | tail -2
echo "=== xpbd helpers as real calls (libnewton_b200_noinline.so)"; NB2_LIB=newton_b200/libnewton_b200_noinline.so timeout -k 5 200 python scripts/quick_bench.py 4096 8 quad xpbd 2>&1 | tail -2
echo "=== featherstone envs=16384"; timeout -k 5 200 python scripts/quick_bench.py 16384 8 quad featherstone 2>&1 | tail -2
} > $O/r2n_batch_sizes.txt 2>&1
cat $O/r2n_batch_sizes.txt
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:featherstone_step_kernel -s 100 -c 1 -f -o $O/r2n_featherstone python scripts/quick_bench.py 4096 8 quad featherstone > $O/r2n_ncu_fs.log 2>&1
tail -1 $O/r2n_ncu_fs.log
python scripts/smoke_entry.py > $O/r2n_smoke.txt 2>&1; tail -2 $O/r2n_smoke.txt
