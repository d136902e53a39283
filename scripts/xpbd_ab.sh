#!/bin/bash
# Round-2 A/B of xpbd_step_kernel variants on one B200: parity (bit-exact suite) + quick_bench timings per variant.
# Usage (under gpurun): bash scripts/xpbd_ab.sh > gpurun_out/r2b_xpbd_ab.txt 2>&1
cd "$(dirname "$0")/.."
AB=newton_b200/libnewton_b200_ab.so
run() {  # label, env assignments...
  local label="$1"; shift
  echo "=== $label"
  env "$@" timeout -k 5 120 python scripts/quick_bench.py 4096 8 quad xpbd 2>&1 | tail -2
}
par() {
  local label="$1"; shift
  echo "--- parity $label"
  env "$@" timeout -k 5 300 python -m pytest tests/test_gpu_xpbd_parity.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -2
}
for W in 1 2 4 7 14; do
  run "warps=$W tma=0" NB2_LIB=$AB NB2_XPBD_WARPS=$W NB2_XPBD_TMA=0
  run "warps=$W tma=1" NB2_LIB=$AB NB2_XPBD_WARPS=$W NB2_XPBD_TMA=1
done
for W in 2 4 7 14; do
  run "warps=$W tma=1 phase_sync" NB2_LIB=$AB NB2_XPBD_WARPS=$W NB2_XPBD_TMA=1 NB2_XPBD_PHASE_SYNC=1
done
run "warps=2 tma=1 no joint cache" NB2_LIB=$AB NB2_XPBD_WARPS=2 NB2_XPBD_NO_JOINT_CACHE=1
for W in 1 2 4 14; do par "warps=$W tma=1" NB2_LIB=$AB NB2_XPBD_WARPS=$W; done
par "warps=2 tma=0" NB2_LIB=$AB NB2_XPBD_WARPS=2 NB2_XPBD_TMA=0
par "warps=4 phase_sync" NB2_LIB=$AB NB2_XPBD_WARPS=4 NB2_XPBD_PHASE_SYNC=1
echo "=== register caps (min resident warps 20 / 24 -> <= 102 / 85 registers), 4096 / 16384 / 32768 envs"
for lib in newton_b200/libnewton_b200_ab.so newton_b200/libnewton_b200_mw20.so newton_b200/libnewton_b200_mw24.so; do
  for E in 4096 16384 32768; do
    for W in 2; do
      echo "--- $lib envs=$E warps=$W"
      NB2_LIB=$lib NB2_XPBD_WARPS=$W timeout -k 5 200 python scripts/quick_bench.py $E 8 quad xpbd 2>&1 | tail -2
    done
  done
done
