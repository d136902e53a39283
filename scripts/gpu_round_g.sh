#!/bin/bash
# round-2g GPU pass: full GPU suite, Featherstone header-cache timing, compute-sanitizer (memcheck + racecheck) over the round-2
# kernels with multi-warp CTAs forced, ncu launch list of the bench, headline bench + reference arm
cd "$(dirname "$0")/.."
O=gpurun_out
timeout -k 5 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > $O/r2g_gpu_tests.txt
cat $O/r2g_gpu_tests.txt
{
echo "=== featherstone (joint headers in smem)"; timeout -k 5 120 python scripts/quick_bench.py 4096 8 quad featherstone 2>&1 | tail -2
echo "=== xpbd"; timeout -k 5 120 python scripts/quick_bench.py 4096 8 quad xpbd 2>&1 | tail -2
} > $O/r2g_kernels.txt 2>&1
cat $O/r2g_kernels.txt
export NB2_XPBD_WARPS=14 NB2_FS_WARPS=14 NB2_COLLIDE_WARPS=8
SEL='quadruped_100 and 33-4 or box_stacks and 1 or convex_pile or hull_pile or restitution'
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_xpbd_parity.py -x -q -k "$SEL" > $O/r2g_sanitize_${tool}_xpbd.log 2>&1
  tail -3 $O/r2g_sanitize_${tool}_xpbd.log
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_featherstone_parity.py tests/test_broad_phase_and_matching.py -m gpu -x -q -k "not full_size" > $O/r2g_sanitize_${tool}_fs_bp.log 2>&1
  tail -3 $O/r2g_sanitize_${tool}_fs_bp.log
done
unset NB2_XPBD_WARPS NB2_FS_WARPS NB2_COLLIDE_WARPS
grep -h "ERROR SUMMARY\|RACECHECK SUMMARY\|passed\|failed" $O/r2g_sanitize_*.log > $O/r2g_sanitize_summary.txt; cat $O/r2g_sanitize_summary.txt
timeout -k 5 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 240 --csv --log-file $O/r2g_launches.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r2g_ncu_bench.log 2>&1
python bench.py --steps 20 --warmup 5 > $O/r2g_bench_n1.json 2> $O/r2g_bench_n1.err; cut -c1-300 $O/r2g_bench_n1.json
python bench.py --impl reference --steps 20 --warmup 5 > $O/r2g_bench_ref.json 2> $O/r2g_bench_ref.err; cut -c1-300 $O/r2g_bench_ref.json
