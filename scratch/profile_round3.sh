#!/bin/bash
# end-of-round-2 profile (2048^2 fp32, default = relaxed engine with the tensor-map pipelines): bench lines, launch list of one bench step,
# one `ncu --set full` capture of each main kernel (summarised ON the box: gpurun returns at most 64 MiB)
TAG=${1:-r2f}
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err || exit 1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
OF2D_MATH=exact python bench.py --no-fp64 --batch 0 > gpurun_out/${TAG}_bench_exact_engine.json 2> /dev/null
python bench.py --steps 1 --warmup 1 --quick > gpurun_out/quick_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 1 --quick > gpurun_out/quick_ncu.log 2>&1
rm -f gpurun_out/*.ncu-rep
cap() {  # method(s) kernel-regex skip count name
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c $4 -f -o gpurun_out/$5 python bench.py --steps 1 --warmup 0 --quick --methods $1 > gpurun_out/ncu_$5.log 2>&1 || echo "ncu $5 failed"
  python scratch/ncu_box.py $TAG 30 > gpurun_out/ncu_box_$5.log 2>&1
}
cap thirion 'k_rt_' 20 2 thirion
cap diffeomorphic 'k_rt_force|k_e_square' 20 3 diffeo
cap elastic 'k_sor_tile' 10 1 elastic
cap fluid 'k_rt_fl|k_fl_|k_sor_tile|k_e_compose' 30 5 fluid
cap curvature 'k_rg_' 10 2 curvature
cap diffusion 'k_hs_pair' 5 1 diffusion
python scratch/launch_summary.py gpurun_out/launches_$TAG.csv > gpurun_out/${TAG}_launches.txt 2>&1
ls gpurun_out/summ | grep $TAG | head -40
