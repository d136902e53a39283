#!/bin/bash
# fused export with the 256-tile look-back window: A/B against the two-kernel path
cd "$(dirname "$0")/.."
O=gpurun_out
{
for i in 1 2; do
echo "=== xpbd, fused export (default)"; timeout -k 5 120 python scripts/quick_bench.py 4096 8 quad xpbd 2>&1 | tail -2
echo "=== xpbd, split (NB2_COLLIDE_FUSED_EXPORT=0)"; NB2_COLLIDE_FUSED_EXPORT=0 timeout -k 5 120 python scripts/quick_bench.py 4096 8 quad xpbd 2>&1 | tail -2
done
echo "=== bench fused"; python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-160
echo "=== bench split"; NB2_COLLIDE_FUSED_EXPORT=0 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-160
for cfg in "4096 quad 12" "3000 stacks 3" "2000 heap 3"; do
  set -- $cfg
  a=$(timeout -k 5 300 python scripts/export_digest.py $1 $2 $3 2>&1 | tail -1)
  b=$(NB2_COLLIDE_FUSED_EXPORT=0 timeout -k 5 300 python scripts/export_digest.py $1 $2 $3 2>&1 | tail -1)
  if [ "$a" == "$b" ] && [ -n "$a" ]; then echo "SAME $a"; else echo "DIFF"; echo "  fused: $a"; echo "  split: $b"; fi
done
timeout -k 5 600 python -m pytest tests/test_gpu_featherstone_tile.py tests/test_broad_phase_and_matching.py tests/test_gpu_xpbd_parity.py -m gpu -q -x 2>&1 | tail -5
} > $O/r2h2_fused_export_ab.txt 2>&1
cat $O/r2h2_fused_export_ab.txt
