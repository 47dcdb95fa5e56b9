run() { echo "logn=$1 c=$2 L=$3: $(TKM_SWEEP_LOGN=$1 TKM_MSM_C=$2 TKM_MSM_TREE_LEVELS=$3 python scripts/msm_sweep.py 2>&1 | tail -1 | cut -c1-140)"; }
run 23 16 5; run 23 16 4; run 23 19 2; run 23 19 3; run 23 19 4
run 24 16 6; run 24 16 5; run 24 19 3; run 24 19 4; run 24 19 5
run 21 16 3; run 21 16 4; run 21 15 4
run 20 16 2; run 20 16 3; run 20 15 3; run 20 14 4; run 20 15 4
run 22 16 4; run 22 16 5; run 22 15 5
