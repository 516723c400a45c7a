#!/bin/bash
# Run on the GPU box (under gpurun): plain bench run first, then the ncu launch list of the SAME command, then one
# `--set full` capture of every kernel of one step (launch set).  Outputs go to gpurun_out/.
# usage: profiles/run_ncu.sh [tag] [launches-per-step] [pairs]     (pairs 1024 = 2048 images = 2.9 GB of pyramid per launch set, >> L2)
set -u
TAG=${1:-r02}
NL=${2:-27}
PAIRS=${3:-1024}
CMD="python bench.py --pairs $PAIRS --steps 1 --warmup 3 --no-e2e --no-cpu --no-latency --no-stereo --no-large"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.json 2> gpurun_out/plain_$TAG.err || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.err; exit 1; }
# launches: 1 synth + 3 warm-up steps + 1 timed step + 2 stage-timing steps, NL launches each
ncu --metrics gpu__time_duration.sum --clock-control none -c $((1 + 5 * NL)) --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
# one whole step (the 4th: after the warm-up) with the full metric set
ncu --set full --clock-control none --import-source on -s $((1 + 3 * NL)) -c $NL -f -o gpurun_out/prof_${TAG} $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log
