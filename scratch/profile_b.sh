#!/bin/bash
python bench.py --steps 1 --warmup 0 --quick > gpurun_out/quick_plain.log 2>&1 || exit 1
for spec in fluid:k_fl_integrate:fluid_integrate fluid:k_fl_increment:fluid_increment fluid:k_sor_tile:sor_tile_fluid elastic:k_sor_tile:sor_tile_elastic diffusion:k_hs_iter:hs_iter; do
  IFS=: read m k t <<< "$spec"
  ncu --set full --clock-control none --import-source on -k regex:$k -s 10 -c 1 -f -o gpurun_out/r1_$t python bench.py --steps 1 --warmup 0 --quick --methods $m > gpurun_out/ncu_$t.log 2>&1 || echo "ncu $t failed"
done
ncu --set full --clock-control none --import-source on -k regex:k_e_conv -s 11 -c 1 -f -o gpurun_out/r1_conv_logger python bench.py --steps 1 --warmup 0 --quick --methods thirion > gpurun_out/ncu_conv.log 2>&1
ls gpurun_out/r1_*.ncu-rep | wc -l
