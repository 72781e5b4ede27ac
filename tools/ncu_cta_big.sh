python tools/prof_solve.py cta LMPC 50 16384 2 > gpurun_out/plain_c2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ipm_kernel_cta -s 1 -c 1 -f -o gpurun_out/r2_cta_big_lmpc_b16384 python tools/prof_solve.py cta LMPC 50 16384 2 > gpurun_out/ncu_c2.log 2>&1
tail -1 gpurun_out/plain_c2.log; ls -la gpurun_out/*.ncu-rep
