#!/bin/bash
# usage: profiles/run_ncu_kernel.sh <kernel-regex> <tag> [skip] [count] [pairs] — plain run, then one --set full capture of the matching kernel(s)
set -u
K=$1; TAG=$2; SKIP=${3:-3}; CNT=${4:-1}; PAIRS=${5:-128}
CMD="python bench.py --pairs $PAIRS --steps 1 --warmup 3 --no-e2e --no-cpu --no-latency --no-stereo --no-large"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.json 2> gpurun_out/plain_$TAG.err || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.err; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c $CNT -f -o gpurun_out/prof_${TAG} $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log
