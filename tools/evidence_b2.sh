# round-2 evidence, ncu part 2: full capture of the stream mapping (N = 50, batch 262144)
python tools/prof_solve.py stream LBMPC 50 262144 2 > gpurun_out/plain_s.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ipm_stream -s 1 -c 1 -f -o gpurun_out/r2_stream_b262144_v3 python tools/prof_solve.py stream LBMPC 50 262144 2 > gpurun_out/ncu_s.log 2>&1
ls -la gpurun_out/*.ncu-rep
