# launch list (kernel durations) of a few Monte-Carlo closed-loop steps at the config-4 width
python tools/loop_bench.py 125000 6 auto > gpurun_out/plain_loop.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2_closed_loop_launch_list.csv python tools/loop_bench.py 125000 6 auto > gpurun_out/ncu_loop.log 2>&1
tail -2 gpurun_out/plain_loop.log
