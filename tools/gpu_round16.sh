#!/bin/bash
mkdir -p gpurun_out
for c in 4 8 16 32 64; do
  echo "== ctas_per_sm $c"
  BLAST_RENDER_CTAS_PER_SM=$c timeout 300 python bench.py --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['kernel_ms'])"
  BLAST_RENDER_CTAS_PER_SM=$c timeout 200 python tools/bench_render.py --skip-c4 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:(v.get('ms'),v.get('GBps')) for k,v in d.items() if 'ms' in v})"
done
