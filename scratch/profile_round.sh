#!/bin/bash
# round profile (2048^2 fp32): launch list of one bench step + one `ncu --set full` capture of each main kernel.
# Reports are summarised ON the box (scratch/ncu_box.py) because gpurun returns at most 64 MiB; every capture is
# summarised as soon as it exists so that a time-out keeps what was done.
TAG=${1:-r1c}
python bench.py --steps 1 --warmup 1 --quick > gpurun_out/quick_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 1 --quick > gpurun_out/quick_ncu.log 2>&1
cap() {  # method(s) kernel-regex skip count name
  timeout 200 ncu --set full --clock-control none -k "regex:$2" -s $3 -c $4 -f -o gpurun_out/$5 python bench.py --steps 1 --warmup 0 --quick --methods $1 > gpurun_out/ncu_$5.log 2>&1 || echo "ncu $5 failed"
  python scratch/ncu_box.py $TAG 30 > gpurun_out/ncu_box_$5.log 2>&1
}
cap fluid 'k_fl_|k_sor_tile|k_e_compose' 30 5 fluid
cap elastic 'k_sor_tile' 10 1 elastic
cap thirion 'k_e_(demons_force|conv|compose)' 40 4 thirion
cap diffusion 'k_hs_iter' 10 1 diffusion
cap curvature 'k_rg_' 30 3 curvature
cap diffeomorphic 'k_e_(conv|square)' 61 2 diffeo
ls -la gpurun_out/summ | head -40
