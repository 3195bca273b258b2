#!/bin/bash
# Round-2 evidence, final capture of the round (fourth session: same kernels, current sources): run from the repository root
# on the GPU box:   bash profiles/tools/capture_r2c.sh
# 1. the bench lines (no profiler), 2. the launch list of the same command, 3. ncu --set full of the kernels.
set -x
O=gpurun_out
python bench.py --steps 20 --warmup 5 > $O/r2c_bench_c2.json 2> $O/r2c_bench_c2.err || exit 1
python bench.py --impl reference --steps 1 --warmup 1 --cpu-trees 64 > $O/r2c_bench_c2_reference_arm.json 2>> $O/r2c_bench_c2.err
for s in c1 c3 c4 c5 gtr jac; do python bench.py --side $s; done > $O/r2c_bench_side_configs.jsonl 2>> $O/r2c_bench_c2.err
python profiles/tools/gmodel_kernel.py 5 > $O/r2c_gmodel_kernel.txt 2>&1
python profiles/tools/nj_cluster.py 200 > $O/r2c_nj_cluster.txt 2>&1
python profiles/tools/nj_cluster.py 128 >> $O/r2c_nj_cluster.txt 2>&1
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/r2c_launches_c2.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-strong > /dev/null 2>&1
$NCU --set full --import-source on -k regex:prune_kernel -s 3 -c 1 -o $O/prof_prune_r2c -f python bench.py --steps 2 --warmup 3 --no-cpu --no-strong > /dev/null 2>&1
$NCU --set full --import-source on -k regex:prune_kernel -s 2 -c 1 -o $O/prof_prune_gmodel_r2c -f python profiles/tools/gmodel_kernel.py 3 --only-new > /dev/null 2>&1
$NCU --set full --import-source on -k regex:nj_hard_lean -s 2 -c 1 -o $O/prof_njhardlean_r2c -f python bench.py --side c3 > /dev/null 2>&1
DODO_NJ_LEAN_W=8 DODO_NJ_CLUSTER=8 $NCU --set full --import-source on -k regex:nj_hard_lean -s 2 -c 1 -o $O/prof_njhardcluster_r2c -f python profiles/tools/nj_cluster.py 200 > /dev/null 2>&1
python profiles/tools/align_time.py > $O/r2c_align_time.txt 2>&1
ls -la $O/*r2c*
