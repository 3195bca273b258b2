#!/bin/bash
# Round-2 evidence in one GPU session (run from the repository root on the GPU box):
#   bash profiles/tools/capture_r2.sh
# 1. the bench line (no profiler), 2. the launch list of the same command, 3. ncu --set full of the kernels.
set -x
O=gpurun_out
python bench.py --steps 10 --warmup 3 > $O/r2_bench_c2.json 2> $O/r2_bench_c2.err || exit 1
python bench.py --impl reference --steps 1 --warmup 1 --cpu-trees 64 > $O/r2_bench_c2_reference_arm.json 2>> $O/r2_bench_c2.err
for s in c1 c3 c4 c5 gtr jac; do python bench.py --side $s; done > $O/r2_bench_side_configs.jsonl 2>> $O/r2_bench_c2.err
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/r2_launches_c2.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-strong > /dev/null 2>&1
$NCU --set full --import-source on -k regex:prune_kernel -s 3 -c 1 -o $O/prof_prune_r2 -f python bench.py --steps 2 --warmup 3 --no-cpu --no-strong > /dev/null 2>&1
$NCU --set full --import-source on -k regex:nj_soft_lean -s 3 -c 1 -o $O/prof_njsoftlean_r2 -f python bench.py --steps 2 --warmup 3 --no-cpu --no-strong > /dev/null 2>&1
$NCU --set full --import-source on -k regex:nj_soft_bwd_lean -s 3 -c 1 -o $O/prof_njsoftbwd_r2 -f python bench.py --steps 2 --warmup 3 --no-cpu --no-strong > /dev/null 2>&1
# the other pruning variants: C3 (large tables, rescaling, JC69, forward only), C5 (+ gradient), learnable GTR (d lnL/d P)
$NCU --set full --import-source on -k regex:prune_kernel -s 2 -c 1 -o $O/prof_prune_c3_r2 -f python bench.py --side c3 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:prune_kernel -s 2 -c 1 -o $O/prof_prune_c5_r2 -f python bench.py --side c5 > /dev/null 2>&1
$NCU --set full --import-source on -k 'regex:prune_kernel<2' -s 1 -c 1 -o $O/prof_prune_gmats_r2 -f python bench.py --side gtr > /dev/null 2>&1
$NCU --set full --import-source on -k regex:blens_jac -s 1 -c 1 -o $O/prof_jac_r2 -f python bench.py --side jac > /dev/null 2>&1
$NCU --set full --import-source on -k regex:gram_logdet -s 1 -c 1 -o $O/prof_logdet_r2 -f python bench.py --side jac > /dev/null 2>&1
$NCU --set full --import-source on -k regex:nj_hard_lean -s 2 -c 1 -o $O/prof_njhardlean_r2 -f python bench.py --side c3 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:geo_soft -c 2 -o $O/prof_geo_r2 -f python -m pytest tests/test_geodesic_gpu.py -m gpu -q -x > /dev/null 2>&1
ls -la $O/*_r2*
