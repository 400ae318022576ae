#!/bin/bash
# knob sweep for the tiled SOR sweep (elastic + fluid at 2048^2): ring depth, CTAs per SM, threads per CTA
run() { echo "== $*"; env "$@" timeout 120 python bench.py --quick --steps 2 --warmup 1 --methods elastic,fluid 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k: round(v,3) for k,v in d['ms'].items()}, d['iterations'])"; }
run OF2D_SOR_NS=6 OF2D_SOR_PER_SM=4
run OF2D_SOR_NS=4 OF2D_SOR_PER_SM=6
run OF2D_SOR_NS=4 OF2D_SOR_PER_SM=7
run OF2D_SOR_NS=3 OF2D_SOR_PER_SM=8
run OF2D_SOR_NS=3 OF2D_SOR_PER_SM=9
run OF2D_SOR_NT=32 OF2D_SOR_NS=6 OF2D_SOR_PER_SM=8
run OF2D_SOR_NT=32 OF2D_SOR_NS=3 OF2D_SOR_PER_SM=16
run OF2D_SOR_NT=32 OF2D_SOR_NS=4 OF2D_SOR_PER_SM=12
echo "== diffusion (hs launch bounds 256,4)"; timeout 120 python bench.py --quick --steps 3 --warmup 2 --methods diffusion,thirion 2>&1 | tail -1
