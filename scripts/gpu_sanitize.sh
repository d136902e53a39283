#!/bin/bash
# compute-sanitizer passes over the parity tests (small scenes): memcheck + racecheck (shared-memory hazards of the
# warp-synchronous kernels) + initcheck.  Outputs: gpurun_out/sanitize_*.log
mkdir -p gpurun_out
SEL='quadruped_100 and 8-8 or shapes_on_plane or box_stacks and 1 or convex_pile or restitution or contact_force'
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_xpbd_parity.py -x -q -k "$SEL" > gpurun_out/sanitize_${tool}_xpbd.log 2>&1
  tail -4 gpurun_out/sanitize_${tool}_xpbd.log
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_featherstone_parity.py -x -q > gpurun_out/sanitize_${tool}_featherstone.log 2>&1
  tail -4 gpurun_out/sanitize_${tool}_featherstone.log
done
grep -c "ERROR SUMMARY" gpurun_out/sanitize_*.log
grep -h "ERROR SUMMARY\|RACECHECK SUMMARY" gpurun_out/sanitize_*.log
