set -x
python scripts/bench_ilqr.py ilqr 16384 200 10 f64 > gpurun_out/p1_ilqr.log 2>&1
python scripts/bench_lqr_tv.py 65536 > gpurun_out/p1_tv.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/p1_ilqr_launches.csv python scripts/bench_ilqr.py ilqr 16384 200 2 f64 > /dev/null 2>&1
REPS=1 ncu --set full --clock-control none --import-source on -k regex:'k_forward_costs|k_forward_commit|k_ilqr_backward_quad' -s 6 -c 3 -o gpurun_out/p1_ilqr_full python scripts/bench_ilqr.py ilqr 16384 200 2 f64 > gpurun_out/p1_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_riccati_t1_tv' -s 2 -c 1 -o gpurun_out/p1_tv_full python scripts/bench_lqr_tv.py 65536 > gpurun_out/p1_ncu2.log 2>&1
cat gpurun_out/p1_ilqr.log gpurun_out/p1_tv.log
