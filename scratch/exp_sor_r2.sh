#!/bin/bash
# relaxed SOR sweep: halo truncation x CTAs per SM -> time (bench --quick) and max |du| vs the 2048^2 fixture
for m in elastic fluid; do
for eps in -30 -26 -22 -18 -14; do
for psm in 4 6 8; do
  t=$(OF2D_SOR_EPS_LOG2=$eps OF2D_SOR_PER_SM=$psm python bench.py --quick --steps 3 --warmup 1 --methods $m 2>/dev/null | python -c "import json,sys; print(round(json.loads(sys.stdin.read().strip().splitlines()[-1])['ms']['$m'],3))")
  d=$(OF2D_SOR_EPS_LOG2=$eps OF2D_SOR_PER_SM=$psm python scratch/parity_full.py $m 2>/dev/null | python -c "
import json,sys
for ln in sys.stdin:
    d=json.loads(ln)
    if d.get('level')=='relaxed': print('du=%.2e iters=%d/%d' % (d['du'], d['iters'], d['ref_iters']))")
  echo "$m eps=$eps per_sm=$psm ms=$t $d"
done; done; done
