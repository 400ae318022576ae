python -m pytest tests/test_solvers_gpu.py tests/test_relaxed_gpu.py -m gpu -q -k "dct2d or curvature" > gpurun_out/r2p_tests.log 2>&1; tail -3 gpurun_out/r2p_tests.log
python scratch/bench_curv_sizes.py > gpurun_out/r2p_curv.log 2>&1; cat gpurun_out/r2p_curv.log
for w in 64 128 256; do OF2D_BENCH_WAVE=$w python bench.py --steps 1 --warmup 1 --no-fp64 --batch 1024 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); b=l['batch']
print('wave', b['config']['wave'], {m:(round(b[m]['resident']['pairs_per_s']), round(b[m]['e2e']['pairs_per_s'])) for m in ('thirion','fluid')})"; done
