for tma in 3; do for kw in 3 5; do for sz in "160 96" "256 256"; do
echo "== tma=$tma kw=$kw size=$sz"; OF2D_FUSED_TMA=$tma timeout 60 python scratch/dbg_kw3.py $sz $kw 2>&1 | tail -1 | cut -c1-200
done; done; done
bash scratch/ab_demons.sh
