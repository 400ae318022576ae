#!/bin/bash
# targeted captures (2048^2 fp32) of the kernels changed last: two-step Horn-Schunck, SOR sweep with the producer warp
TAG=${1:-r1d}
cap() {  # method(s) kernel-regex skip count name
  timeout 200 ncu --set full --clock-control none -k "regex:$2" -s $3 -c $4 -f -o gpurun_out/$5 python bench.py --steps 1 --warmup 0 --quick --methods $1 > gpurun_out/ncu_$5.log 2>&1 || echo "ncu $5 failed"
}
cap diffusion 'k_hs_pair' 5 1 diffusion
cap elastic 'k_sor_tile' 10 1 elastic
cap fluid 'k_sor_tile' 10 1 fluid
python scratch/ncu_box.py $TAG 40 > gpurun_out/ncu_box.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 1 --quick > gpurun_out/quick_ncu.log 2>&1
ls gpurun_out/summ
