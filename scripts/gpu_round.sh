#!/bin/bash
# One GPU session: parity tests, the three single-GPU workloads, ncu evidence.  Outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
TAG=${1:-r1c}
python bench.py --steps 300 --warmup 10 > gpurun_out/${TAG}_bench_quadruped_xpbd.json 2> gpurun_out/${TAG}_bench_q.err
python bench.py --steps 300 --warmup 10 --workload box_stacks_xpbd > gpurun_out/${TAG}_bench_box_stacks_xpbd.json 2> gpurun_out/${TAG}_bench_s.err
python bench.py --steps 200 --warmup 10 --workload quadruped_featherstone > gpurun_out/${TAG}_bench_quadruped_featherstone.json 2> gpurun_out/${TAG}_bench_f.err
tail -c 600 gpurun_out/${TAG}_bench_*.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches_featherstone.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload quadruped_featherstone > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches_box_stacks.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload box_stacks_xpbd > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:featherstone_step -s 30 -c 1 -o gpurun_out/${TAG}_featherstone python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload quadruped_featherstone > gpurun_out/ncu_fs.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:collide_kernel -s 30 -c 1 -o gpurun_out/${TAG}_collide_stacks python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload box_stacks_xpbd > gpurun_out/ncu_cs.log 2>&1
ls -la gpurun_out
