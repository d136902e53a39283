#!/bin/bash
# round-2h GPU pass: fused contact export (collide_kernel<.., EXPORT=true>) - digests vs the two-kernel path at several batch sizes /
# CTA shapes, A/B timing, the whole GPU suite, sanitizer on the collide tests, sticky / report matching, FMA-contracted twin timing, bench
cd "$(dirname "$0")/.."
O=gpurun_out
{
for cfg in "100 quad" "4096 quad" "20000 quad" "512 stacks" "3000 stacks" "64 heap" "2000 heap"; do
  set -- $cfg
  for W in 0 1; do
    a=$(NB2_COLLIDE_WARPS=$W NB2_COLLIDE_FUSED_EXPORT=1 timeout -k 5 300 python scripts/export_digest.py $1 $2 2>&1 | tail -1)
    b=$(NB2_COLLIDE_WARPS=$W NB2_COLLIDE_FUSED_EXPORT=0 timeout -k 5 300 python scripts/export_digest.py $1 $2 2>&1 | tail -1)
    if [ "$a" == "$b" ] && [ -n "$a" ]; then echo "SAME  warps=$W $a"; else echo "DIFF  warps=$W"; echo "  fused: $a"; echo "  split: $b"; fi
  done
done
} > $O/r2h_export_digest.txt 2>&1
cat $O/r2h_export_digest.txt
{
echo "=== xpbd, fused export (default)"; timeout -k 5 120 python scripts/quick_bench.py 4096 8 quad xpbd 2>&1 | tail -2
echo "=== xpbd, collide + contact_export_kernel (NB2_COLLIDE_FUSED_EXPORT=0)"; NB2_COLLIDE_FUSED_EXPORT=0 timeout -k 5 120 python scripts/quick_bench.py 4096 8 quad xpbd 2>&1 | tail -2
echo "=== xpbd, FMA-contracted twin library (not the product: velocities leave the 1e-5 tolerance)"; NB2_LIB=newton_b200/libnewton_b200_fast.so timeout -k 5 120 python scripts/quick_bench.py 4096 8 quad xpbd 2>&1 | tail -2
echo "=== box stacks, fused"; timeout -k 5 120 python scripts/quick_bench.py 512 8 stacks xpbd 2>&1 | tail -2
echo "=== box stacks, split"; NB2_COLLIDE_FUSED_EXPORT=0 timeout -k 5 120 python scripts/quick_bench.py 512 8 stacks xpbd 2>&1 | tail -2
echo "=== featherstone, fused"; timeout -k 5 120 python scripts/quick_bench.py 4096 8 quad featherstone 2>&1 | tail -2
echo "=== 32768 envs, fused"; timeout -k 5 200 python scripts/quick_bench.py 32768 8 quad xpbd 2>&1 | tail -2
} > $O/r2h_kernels.txt 2>&1
cat $O/r2h_kernels.txt
timeout -k 5 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 > $O/r2h_gpu_tests.txt
cat $O/r2h_gpu_tests.txt
for tool in memcheck racecheck; do
  NB2_COLLIDE_WARPS=8 timeout 600 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_xpbd_parity.py -x -q -k "single_collide or deterministic_export or box_stacks and 7" > $O/r2h_sanitize_${tool}.log 2>&1
  tail -2 $O/r2h_sanitize_${tool}.log
done
grep -h "ERROR SUMMARY\|RACECHECK SUMMARY\|passed\|failed" $O/r2h_sanitize_*.log > $O/r2h_sanitize_summary.txt
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2h_bench_n1.json 2> $O/r2h_bench_n1.err; cut -c1-400 $O/r2h_bench_n1.json
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-export-contacts > $O/r2h_bench_n1_no_export.json 2> $O/r2h_bench_n1_no_export.err; cut -c1-200 $O/r2h_bench_n1_no_export.json
