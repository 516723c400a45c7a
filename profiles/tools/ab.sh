#!/bin/bash
# usage (on the GPU box): profiles/tools/ab.sh v1 v2 ...  — per variant library: quick parity check, then per-stage timing
for v in "$@"; do
  cp variants/liborbx_$v.so tc2li_slam_b200/lib/liborbx.so
  echo "== $v"
  python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -1
  profiles/tools/stages.sh
  profiles/tools/stages.sh
done
