#!/bin/bash
# usage: profile_one.sh <method> <kernel-regex> <skip> <name>: one ncu --set full capture (with source) inside bench.py --quick
timeout 300 ncu --set full --import-source on --clock-control none -k "regex:$2" -s $3 -c 1 -f -o gpurun_out/$4 python bench.py --steps 1 --warmup 0 --quick --methods $1 > gpurun_out/ncu_$4.log 2>&1 || echo "ncu $4 failed"
ls -la gpurun_out/$4.ncu-rep
