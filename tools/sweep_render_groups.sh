#!/bin/bash
# Sweep K4's work-item granularity (BLAST_RENDER_CTAS_PER_SM x BLAST_RENDER_MIN_GROUP) on the C2-mix and C3 shapes.
#   gpurun --timeout 900 -- 'bash tools/sweep_render_groups.sh'
mkdir -p gpurun_out
out=gpurun_out/sweep_render_groups.txt
: > $out
for cfg in "32 64" "64 32" "128 32" "16 64"; do
  set -- $cfg
  for shape in "1024 720000" "4096 1048576"; do
    set -- $cfg $shape
    r=$(BLAST_RENDER_CTAS_PER_SM=$1 BLAST_RENDER_MIN_GROUP=$2 timeout 60 python tools/bench_render.py --voices $3 --frames $4 --skip-c4 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print(' '.join('%s=%.4f' % (k, v['ms']) for k, v in d.items()))")
    echo "ctas_per_sm=$1 min_group=$2 voices=$3 frames=$4: $r" | tee -a $out
  done
done
