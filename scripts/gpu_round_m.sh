#!/bin/bash
# round-2m: fixed heap test, compute-sanitizer (memcheck + racecheck) over the kernels touched since round 2g (collide staging write-out,
# speculative collide / broad phase, D6 two-axis Featherstone), launch list + ncu --set full of the final collide / xpbd kernels, bench lines
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout -k 5 600 python -m pytest tests/test_speculative_contacts.py -m gpu -q -x -k heap 2>&1 | tail -5 > $O/r2m_heap_test.txt; cat $O/r2m_heap_test.txt
export NB2_COLLIDE_WARPS=8
SEL='quadruped_100 and 33-4 or box_stacks and 1 or convex_pile or hull_pile'
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_xpbd_parity.py -x -q -k "$SEL" > $O/r2m_sanitize_${tool}_xpbd.log 2>&1
  tail -2 $O/r2m_sanitize_${tool}_xpbd.log
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_speculative_contacts.py tests/test_d6_two_angular_axes.py tests/test_broad_phase_and_matching.py -m gpu -x -q -k "not contact_report" > $O/r2m_sanitize_${tool}_spec.log 2>&1
  tail -2 $O/r2m_sanitize_${tool}_spec.log
done
unset NB2_COLLIDE_WARPS
grep -h "ERROR SUMMARY\|RACECHECK SUMMARY\|passed\|failed" $O/r2m_sanitize_*.log > $O/r2m_sanitize_summary.txt; cat $O/r2m_sanitize_summary.txt
timeout -k 5 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 240 --csv --log-file $O/r2m_launches.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r2m_ncu_bench.log 2>&1
for k in collide_kernel xpbd_step_kernel; do
  timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 200 -c 1 -f -o $O/r2m_$k python scripts/quick_bench.py 4096 8 quad xpbd > $O/r2m_ncu_$k.log 2>&1
  tail -1 $O/r2m_ncu_$k.log
done
python bench.py --steps 30 --warmup 5 > $O/r2m_bench_n1.json 2> $O/r2m_bench_n1.err; cut -c1-200 $O/r2m_bench_n1.json
python bench.py --impl reference --steps 20 --warmup 5 > $O/r2m_bench_ref.json 2> $O/r2m_bench_ref.err; cut -c1-300 $O/r2m_bench_ref.json
for w in box_stacks_xpbd quadruped_featherstone quadruped_xpbd_stock; do
  python bench.py --workload $w --steps 30 --warmup 5 --no-cpu-baseline > $O/r2m_bench_$w.json 2>/dev/null; cut -c1-160 $O/r2m_bench_$w.json
done
