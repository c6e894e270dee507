#!/bin/bash
# Times the Criteo-shaped step with parts of fused_rows_kernel switched off (DFM_FR_ABLATE; results are wrong, timing only).
out=gpurun_out/${1:-fr_ablate}.txt
: > $out
for ab in ${ABL:-0 1 2 4 8 16 32 48 56 64 127}; do
  DFM_FR_ABLATE=$ab timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary --no-zipf 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ablate %3d  step %.4f ms  kernel %.4f ms' % ($ab, d['ms_per_step'], d['phases_ms']['gather']))
" >> $out
done
cat $out
