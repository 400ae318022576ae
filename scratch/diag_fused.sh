for cfg in "1 0" "1 1" "2 0" "3 0" "0 0"; do set -- $cfg; echo "== FUSED=$1 NOFAST=$2"; OF2D_FUSED=$1 OF2D_FUSED_NOFAST=$2 python scratch/diag_diffeo.py diffeomorphic 2048 2>&1 | cut -c1-200; done
