python tools/prof_solve.py stream LBMPC 50 262144 2 > gpurun_out/plain_s2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ipm_stream -s 1 -c 1 -o gpurun_out/r2_stream_b262144_v2 -f python tools/prof_solve.py stream LBMPC 50 262144 2 > gpurun_out/ncu_s2.log 2>&1
tail -2 gpurun_out/plain_s2.log; tail -3 gpurun_out/ncu_s2.log
