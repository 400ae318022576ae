#!/bin/bash
# round-2 targeted captures (2048^2 fp32, default = relaxed engine): the four Thirion kernels, both SOR sweeps, the curvature passes
TAG=${1:-r2a}
WHAT=${2:-thirion,elastic,fluid}
cap() {  # method(s) kernel-regex skip count name
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c $4 -f -o gpurun_out/$5 python bench.py --steps 1 --warmup 0 --quick --methods $1 > gpurun_out/ncu_$5.log 2>&1 || echo "ncu $5 failed"
}
[[ $WHAT == *thirion* ]] && cap thirion 'k_e_(demons_force|conv|compose)|k_rx_' 20 2 thirion
[[ $WHAT == *elastic* ]] && cap elastic 'k_sor_tile' 10 1 elastic
[[ $WHAT == *fluid* ]] && cap fluid 'k_sor_tile|k_fl_integrate' 20 2 fluid
[[ $WHAT == *curvature* ]] && cap curvature 'k_rg_' 30 3 curvature
[[ $WHAT == *diffeo* ]] && cap diffeomorphic 'k_e_(conv|square)' 61 2 diffeo
[[ $WHAT == *diffusion* ]] && cap diffusion 'k_hs_pair' 5 1 diffusion
python scratch/ncu_box.py $TAG 30 > gpurun_out/ncu_box.log 2>&1
ls gpurun_out/summ
