set -x
python bench.py > gpurun_out/bench21.json 2> gpurun_out/bench21.err
tail -c 600 gpurun_out/bench21.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_msm_r01b.csv python bench.py --steps 2 --warmup 3 --skip-aux --skip-prove > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name regex:k_accumulate --launch-skip 3 --launch-count 1 -f -o gpurun_out/k_accumulate_b python bench.py --steps 1 --warmup 3 --skip-aux --skip-prove > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name regex:k_final --launch-skip 3 --launch-count 1 -f -o gpurun_out/k_final_b python bench.py --steps 1 --warmup 3 --skip-aux --skip-prove > gpurun_out/ncu_f.log 2>&1
python scripts/prove_full.py --fixed-base-tables --warmup 1 --repeats 5 > gpurun_out/prove_full9.json 2> gpurun_out/prove_full9.err
tail -c 300 gpurun_out/prove_full9.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_prove_b.csv python scripts/prove_full.py --repeats 1 --warmup 1 --no-verify > gpurun_out/ncu_p.log 2>&1
ls -la gpurun_out | tail -12
