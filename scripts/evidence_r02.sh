# Final N=1 evidence pass of round 2 (run under gpurun; every ncu command repeats a command that already exited 0 without ncu).
set -x
python bench.py --steps 3 --warmup 3 --skip-aux --skip-prove > gpurun_out/ev_bench.json 2> gpurun_out/ev_bench.err || exit 1
rm -f gpurun_out/*.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_msm_2p22.csv python bench.py --steps 2 --warmup 3 --skip-aux --skip-prove > gpurun_out/ncu_l.log 2>&1
# the pair tree's level-0 forward/apply launches and the XYZZ pass of one timed MSM (after the warm-up MSMs: 4 levels x 2 halves = 8 launches each per MSM)
ncu --set full --clock-control none --import-source on --kernel-name regex:k_tree_apply --launch-skip 24 --launch-count 2 -f -o gpurun_out/r02_k_tree_apply python bench.py --steps 1 --warmup 3 --skip-aux --skip-prove > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name regex:k_tree_fwd --launch-skip 24 --launch-count 2 -f -o gpurun_out/r02_k_tree_fwd python bench.py --steps 1 --warmup 3 --skip-aux --skip-prove > gpurun_out/ncu_f.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name regex:k_accumulate --launch-skip 3 --launch-count 1 -f -o gpurun_out/r02_k_accumulate python bench.py --steps 1 --warmup 3 --skip-aux --skip-prove > gpurun_out/ncu_x.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_prove.csv python scripts/prove_full.py --repeats 1 --warmup 1 --no-verify > gpurun_out/ncu_p.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_bintt_16384x512.csv python scripts/ntt_probe.py > gpurun_out/ncu_n.log 2>&1
# summaries are made on the box; the .ncu-rep files stay there (gpurun_out/ is capped at 64 MiB)
for k in k_tree_apply k_tree_fwd k_accumulate; do python scripts/ncu_summary.py gpurun_out/r02_$k.ncu-rep > gpurun_out/r02_ncu_${k}_summary.csv; rm -f gpurun_out/r02_$k.ncu-rep; done
python scripts/launch_summary.py gpurun_out/r02_launches_prove.csv "setup + 2 proves (1 warm-up, 1 timed) under ncu" > gpurun_out/r02_launches_prove_summary.csv; rm -f gpurun_out/r02_launches_prove.csv
ls -la gpurun_out | tail -14
