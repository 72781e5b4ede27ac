# round-2 evidence, ncu (one GPU): launch list of the headline bench, full captures of the dominant kernels
python bench.py --steps 3 --warmup 3 > gpurun_out/plain_b.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_bench_launch_list.csv python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_ll.log 2>&1
python tools/prof_solve.py warp LBMPC 50 1024 2 > gpurun_out/plain_w.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ipm_kernel -s 1 -c 1 -f -o gpurun_out/r2_warp_b1024 python tools/prof_solve.py warp LBMPC 50 1024 2 > gpurun_out/ncu_w.log 2>&1
python tools/prof_solve.py stream LBMPC 50 262144 2 > gpurun_out/plain_s.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ipm_stream -s 1 -c 1 -f -o gpurun_out/r2_stream_b262144_v3 python tools/prof_solve.py stream LBMPC 50 262144 2 > gpurun_out/ncu_s.log 2>&1
python tools/prof_solve.py auto LBMPC 200 65536 2 > gpurun_out/plain_n200.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ipm_ -s 2 -c 2 -f -o gpurun_out/r2_n200_b65536_v2 python tools/prof_solve.py auto LBMPC 200 65536 2 > gpurun_out/ncu_n200.log 2>&1
cat gpurun_out/plain_*.log | tail -5; ls -la gpurun_out/*.ncu-rep
