#!/bin/bash
python bench.py --steps 1 --warmup 0 --quick --methods curvature > gpurun_out/ncu_plain.log 2>&1 || exit 1
for spec in curvature:k_rg_cols:curv_cols curvature:k_rg_rows_fwd:curv_rows_fwd curvature:k_rg_rows_inv:curv_rows_inv; do
  IFS=: read m k t <<< "$spec"
  ncu --set full --clock-control none --import-source on -k regex:$k -s 10 -c 1 -f -o gpurun_out/r1b_$t python bench.py --steps 1 --warmup 0 --quick --methods $m > gpurun_out/ncu_$t.log 2>&1 || echo "ncu $t failed"
done
ls gpurun_out/*.ncu-rep
