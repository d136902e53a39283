#!/bin/bash
# round-2j (re-entry): whole GPU suite, fused-export A/B (results of round h2 were lost with the container), unrolled-row variant, bench lines
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout -k 5 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 > $O/r2j_gpu_tests.txt
tail -3 $O/r2j_gpu_tests.txt
{
for i in 1 2; do
echo "=== xpbd, fused export (default)"; timeout -k 5 120 python scripts/quick_bench.py 4096 8 quad xpbd 2>&1 | tail -2
echo "=== xpbd, split (NB2_COLLIDE_FUSED_EXPORT=0)"; NB2_COLLIDE_FUSED_EXPORT=0 timeout -k 5 120 python scripts/quick_bench.py 4096 8 quad xpbd 2>&1 | tail -2
echo "=== xpbd, unrolled joint rows"; NB2_LIB=newton_b200/libnewton_b200_unroll.so timeout -k 5 120 python scripts/quick_bench.py 4096 8 quad xpbd 2>&1 | tail -2
done
echo "=== featherstone"; timeout -k 5 120 python scripts/quick_bench.py 4096 8 quad featherstone 2>&1 | tail -2
echo "=== box stacks"; timeout -k 5 120 python scripts/quick_bench.py 512 8 stacks xpbd 2>&1 | tail -2
for cfg in "4096 quad 12" "3000 stacks 3" "2000 heap 3"; do
  set -- $cfg
  a=$(timeout -k 5 300 python scripts/export_digest.py $1 $2 $3 2>&1 | tail -1)
  b=$(NB2_COLLIDE_FUSED_EXPORT=0 timeout -k 5 300 python scripts/export_digest.py $1 $2 $3 2>&1 | tail -1)
  if [ "$a" == "$b" ] && [ -n "$a" ]; then echo "SAME $a"; else echo "DIFF"; echo "  fused: $a"; echo "  split: $b"; fi
done
} > $O/r2j_fused_export_ab.txt 2>&1
cat $O/r2j_fused_export_ab.txt
python bench.py --steps 30 --warmup 5 > $O/r2j_bench_n1.json 2> $O/r2j_bench_n1.err; cut -c1-600 $O/r2j_bench_n1.json
NB2_COLLIDE_FUSED_EXPORT=0 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/r2j_bench_n1_split.json 2>/dev/null; cut -c1-200 $O/r2j_bench_n1_split.json
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-export-contacts > $O/r2j_bench_n1_no_export.json 2>/dev/null; cut -c1-200 $O/r2j_bench_n1_no_export.json
