#!/bin/bash
python bench.py --steps 1 --warmup 1 --quick > gpurun_out/quick_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 1 --quick > gpurun_out/quick_ncu.log 2>&1
for spec in curvature:k_cf_cols:curv_cols curvature:k_cf_rows_fwd:curv_rows_fwd curvature:k_cf_rows_inv:curv_rows_inv thirion:k_e_demons_force:demons_force thirion:k_e_compose:compose; do
  IFS=: read m k t <<< "$spec"
  ncu --set full --clock-control none --import-source on -k regex:$k -s 10 -c 1 -f -o gpurun_out/r1_$t python bench.py --steps 1 --warmup 0 --quick --methods $m > gpurun_out/ncu_$t.log 2>&1 || echo "ncu $t failed"
done
ls gpurun_out/r1_*.ncu-rep | wc -l
