#!/bin/bash
# one full ncu capture per top kernel (steady-state launch #10 of its method); plain run first
set -e
python bench.py --steps 1 --warmup 0 --quick > gpurun_out/ncu_plain.log 2>&1
run() {  # method kernel-regex tag
  ncu --set full --clock-control none --import-source on -k regex:$2 -s 10 -c 1 -f -o gpurun_out/prof_$3 \
      python bench.py --steps 1 --warmup 0 --quick --methods $1 > gpurun_out/ncu_$3.log 2>&1 || echo "ncu $3 failed"
}
run curvature k_curv_cols curv_cols
run curvature k_curv_rows_fwd curv_rows_fwd
run thirion "k_e_conv" conv
run thirion k_e_demons_force demons_force
run fluid k_fl_integrate fl_integrate
run fluid k_fl_increment fl_increment
run fluid k_sor_tile sor_fluid
run elastic k_sor_tile sor_elastic
run diffusion k_hs_iter hs
ls -la gpurun_out/*.ncu-rep
