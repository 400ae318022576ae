#!/bin/bash
# last refresh of round 2 (packed / XOR-addressed register DCT, three-stage Fluid integrate): the full GPU test suite, then scratch/profile_round3.sh
TAG=${1:-r2h}
python -m pytest tests -m gpu -q > gpurun_out/${TAG}_tests.log 2>&1; tail -3 gpurun_out/${TAG}_tests.log
bash scratch/profile_round3.sh $TAG
