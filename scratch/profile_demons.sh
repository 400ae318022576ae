#!/bin/bash
# ncu --set full captures (with source) of the two fused Demons kernels, 2048^2 fp32; reports come back in gpurun_out/
cap() {
  timeout 300 ncu --set full --import-source on --clock-control none -k "regex:$2" -s $3 -c $4 -f -o gpurun_out/$5 python bench.py --steps 1 --warmup 0 --quick --methods $1 > gpurun_out/ncu_$5.log 2>&1 || echo "ncu $5 failed"
}
cap thirion 'k_rt_compose_conv' 6 1 r2z_compose
cap thirion 'k_rt_force_conv' 6 1 r2z_force
ls -la gpurun_out/*.ncu-rep
