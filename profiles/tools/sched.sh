#!/bin/bash
# per scheduling variant of the blur (ORBX_SCHED, orbx_host.cu): whole-step time of the resident run
for v in 0 1 2 3 4; do echo "== sched $v"; ORBX_SCHED=$v profiles/tools/stages.sh; ORBX_SCHED=$v profiles/tools/stages.sh; done
