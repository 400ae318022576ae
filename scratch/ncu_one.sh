#!/bin/bash
# usage: ncu_one.sh method kernel-regex tag [skip]
set -e
python bench.py --steps 1 --warmup 0 --quick --methods $1 > gpurun_out/ncu_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:$2 -s ${4:-10} -c 1 -f -o gpurun_out/prof_$3 python bench.py --steps 1 --warmup 0 --quick --methods $1 > gpurun_out/ncu_$3.log 2>&1 || echo "ncu $3 failed"
