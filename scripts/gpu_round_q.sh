#!/bin/bash
# round-2q (final of round 2): GPU suite, memcheck over the kernels touched in 2o-2q (Featherstone contact pass / H tables / Cholesky,
# collide segments + mesh-plane walk), launch list of the headline bench, bench lines of all four configs + the reference arm
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout -k 5 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > $O/r2q_gpu_tests.txt; tail -3 $O/r2q_gpu_tests.txt
NB2_COLLIDE_WARPS=8 timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_mesh_plane.py tests/test_gpu_featherstone_parity.py -m gpu -x -q -k "mesh_plane_matches or quadruped" > $O/r2q_sanitize_memcheck.log 2>&1
grep -h "ERROR SUMMARY\|passed\|failed" $O/r2q_sanitize_memcheck.log | tail -3 > $O/r2q_sanitize_summary.txt; cat $O/r2q_sanitize_summary.txt
python bench.py --steps 200 --warmup 10 > $O/r2q_bench_n1.json 2> $O/r2q_bench_n1.err; cut -c1-220 $O/r2q_bench_n1.json; tail -2 $O/r2q_bench_n1.err
python bench.py --impl reference --steps 20 --warmup 5 > $O/r2q_bench_ref.json 2> $O/r2q_bench_ref.err; cut -c1-200 $O/r2q_bench_ref.json
for w in box_stacks_xpbd quadruped_featherstone quadruped_xpbd_stock; do
  python bench.py --workload $w --steps 100 --warmup 5 --no-cpu-baseline > $O/r2q_bench_$w.json 2>/dev/null; cut -c1-160 $O/r2q_bench_$w.json
done
timeout -k 5 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 240 --csv --log-file $O/r2q_launches.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-fast-twin > $O/r2q_ncu_bench.log 2>&1
{ echo "=== featherstone, shuffle substitutions (default)"; python scripts/quick_bench.py 4096 8 quad featherstone 2>&1 | tail -2
  echo "=== featherstone, NB2_FS_SHFL_SUBST=0 (round-2p substitutions)"; NB2_FS_SHFL_SUBST=0 python scripts/quick_bench.py 4096 8 quad featherstone 2>&1 | tail -2; } > $O/r2q_featherstone_subst_ab.txt 2>&1; cat $O/r2q_featherstone_subst_ab.txt
python scripts/smoke_entry.py > $O/r2q_smoke.txt 2>&1; tail -1 $O/r2q_smoke.txt
