#!/bin/bash
# usage: ncu_multi.sh "method:regex:tag" ...
python bench.py --steps 1 --warmup 0 --quick > gpurun_out/ncu_plain.log 2>&1 || exit 1
for spec in "$@"; do
  IFS=: read m k t <<< "$spec"
  ncu --set full --clock-control none --import-source on -k regex:$k -s 10 -c 1 -f -o gpurun_out/prof_$t python bench.py --steps 1 --warmup 0 --quick --methods $m > gpurun_out/ncu_$t.log 2>&1 || echo "ncu $t failed"
done
