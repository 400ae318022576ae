#!/bin/bash
# refresh of the bench lines and of the two Demons captures after the last kernel change; full GPU test suite
TAG=${1:-r2f}
python -m pytest tests -m gpu -q > gpurun_out/${TAG}_tests.log 2>&1; tail -3 gpurun_out/${TAG}_tests.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -2 gpurun_out/${TAG}_bench.err
rm -f gpurun_out/*.ncu-rep
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:k_rt_" -s 20 -c 2 -f -o gpurun_out/thirion python bench.py --steps 1 --warmup 0 --quick --methods thirion > gpurun_out/ncu_thirion.log 2>&1
python scratch/ncu_box.py $TAG 30 > gpurun_out/ncu_box_thirion.log 2>&1
ls gpurun_out/summ | grep ${TAG}_thirion
