# ncu --set full of the stream solve of one Monte-Carlo closed-loop step at the config-4 width (the 4th stream launch of the run)
python tools/loop_bench.py 125000 5 auto > gpurun_out/plain_c4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ipm_stream -s 3 -c 1 -f -o gpurun_out/r2_stream_closed_loop_125k python tools/loop_bench.py 125000 5 auto > gpurun_out/ncu_c4.log 2>&1
tail -1 gpurun_out/plain_c4.log; ls -la gpurun_out/*.ncu-rep
