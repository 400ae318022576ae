#!/bin/bash
# timing of build variants: scratch/libof2d_cuda_*.so copied over the library one after the other
M=${1:-thirion,diffeomorphic}
cp opticalflow2d_b200/lib/libof2d_cuda.so /tmp/orig.so
for f in scratch/libof2d_cuda_*.so; do
  cp $f opticalflow2d_b200/lib/libof2d_cuda.so
  echo "== $f"; python bench.py --quick --steps 5 --warmup 2 --methods $M 2>&1 | tail -1
done
cp /tmp/orig.so opticalflow2d_b200/lib/libof2d_cuda.so
