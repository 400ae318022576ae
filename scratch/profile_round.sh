#!/bin/bash
# round profile (2048^2 fp32): launch list of one bench step + one `ncu --set full` capture of each main kernel.
# Reports are summarised ON the box (scratch/ncu_box.py) because gpurun returns at most 64 MiB.
TAG=${1:-r1b}
python bench.py --steps 1 --warmup 1 --quick > gpurun_out/quick_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 1 --quick > gpurun_out/quick_ncu.log 2>&1
cap() {  # method(s) kernel-regex skip count name
  ncu --set full --clock-control none -k "regex:$2" -s $3 -c $4 -f -o gpurun_out/$5 python bench.py --steps 1 --warmup 0 --quick --methods $1 > gpurun_out/ncu_$5.log 2>&1 || echo "ncu $5 failed"
}
cap thirion 'k_e_(demons_force|conv|compose)' 40 4 thirion
cap diffeomorphic 'k_e_(conv|square)' 61 3 diffeo          # conv_maxabs + the first squarings of an iteration
cap fluid 'k_fl_|k_sor_tile' 30 3 fluid
cap curvature 'k_rg_' 30 3 curvature
cap diffusion 'k_hs_iter' 10 1 diffusion
cap elastic 'k_sor_tile' 10 1 elastic
python scratch/ncu_box.py $TAG 30
ls -la gpurun_out/summ | head -40
