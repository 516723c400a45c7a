#!/bin/bash
# usage: profiles/tools/stages.sh [bench args...]  — prints per-stage ms of one bench run
python bench.py --no-cpu --no-e2e --no-latency --no-stereo --no-large "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],2), {k:round(v['ms'],2) for k,v in d['stages'].items()})"
