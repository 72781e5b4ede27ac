python tools/prof_solve.py auto LMPC 50 16384 2 > gpurun_out/plain_c2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ipm_kernel -s 1 -c 1 -f -o gpurun_out/r2_warp_lmpc_b16384 python tools/prof_solve.py auto LMPC 50 16384 2 > gpurun_out/ncu_c2.log 2>&1
tail -1 gpurun_out/plain_c2.log; ls -la gpurun_out/*.ncu-rep
