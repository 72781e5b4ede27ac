# round-2 evidence, ncu part 3: the two launches of the N = 200 solve (stream mapping + hand-over to the CTA mapping), batch 65536
python tools/prof_solve.py auto LBMPC 200 65536 2 > gpurun_out/plain_n200.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ipm_ -s 2 -c 2 -f -o gpurun_out/r2_n200_b65536_v2 python tools/prof_solve.py auto LBMPC 200 65536 2 > gpurun_out/ncu_n200.log 2>&1
ls -la gpurun_out/*.ncu-rep
