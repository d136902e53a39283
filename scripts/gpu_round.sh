#!/bin/bash
# One GPU session: parity tests, the single-GPU workloads, ncu evidence.  Outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
TAG=${1:-r1f}
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 500 --warmup 10 > gpurun_out/${TAG}_bench_quadruped_xpbd.json 2> gpurun_out/${TAG}_bench_q.err
python bench.py --steps 300 --warmup 10 --workload box_stacks_xpbd > gpurun_out/${TAG}_bench_box_stacks_xpbd.json 2> gpurun_out/${TAG}_bench_s.err
python bench.py --steps 300 --warmup 10 --workload quadruped_featherstone > gpurun_out/${TAG}_bench_quadruped_featherstone.json 2> gpurun_out/${TAG}_bench_f.err
python bench.py --steps 300 --warmup 10 --workload quadruped_xpbd_stock --no-cpu-baseline > gpurun_out/${TAG}_bench_quadruped_xpbd_stock.json 2> gpurun_out/${TAG}_bench_k.err
python bench.py --steps 300 --warmup 10 --fast-fp --no-cpu-baseline > gpurun_out/${TAG}_bench_quadruped_xpbd_fastfp.json 2> gpurun_out/${TAG}_bench_ff.err
python bench.py --steps 200 --warmup 10 --envs 16384 --no-cpu-baseline > gpurun_out/${TAG}_bench_quadruped_xpbd_16k.json 2> gpurun_out/${TAG}_bench_16k.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2> gpurun_out/${TAG}_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_quadruped_xpbd.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_featherstone.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload quadruped_featherstone > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_box_stacks.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload box_stacks_xpbd > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:xpbd_step -s 30 -c 1 -o gpurun_out/${TAG}_xpbd python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_x.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:featherstone_step -s 30 -c 1 -o gpurun_out/${TAG}_featherstone python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload quadruped_featherstone > gpurun_out/ncu_f.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:collide_kernel -s 30 -c 1 -o gpurun_out/${TAG}_collide_quad python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c.log 2>&1
python __graft_entry__.py --smoke 2>&1 | tail -1
ls -la gpurun_out | tail -20
