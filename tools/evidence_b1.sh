# round-2 evidence, ncu part 1: launch list of the headline bench + full capture of the warp mapping at the headline shape
python bench.py --steps 3 --warmup 3 > gpurun_out/plain_b.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_bench_launch_list.csv python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_ll.log 2>&1
python tools/prof_solve.py warp LBMPC 50 1024 2 > gpurun_out/plain_w.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ipm_kernel -s 1 -c 1 -f -o gpurun_out/r2_warp_b1024 python tools/prof_solve.py warp LBMPC 50 1024 2 > gpurun_out/ncu_w.log 2>&1
ls -la gpurun_out/*.ncu-rep
