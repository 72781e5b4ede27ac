# round-2 evidence, one GPU: full GPU test suite, all bench configs (+ reference arm), mapping sweep
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r2_gputest.log; cat gpurun_out/r2_gputest.log
python bench.py --impl reference > gpurun_out/r2_bench_reference_arm.json 2>/dev/null
for c in 1 2 3 4; do python bench.py --config $c $( [ $c = 4 ] && echo "--steps 2" ) > gpurun_out/r2_bench_c$c.json 2> gpurun_out/r2_bench_c$c.err; tail -c 400 gpurun_out/r2_bench_c$c.err; done
python tools/kernel_sweep.py r2_kernel_sweep.json > gpurun_out/r2_kernel_sweep.log 2>&1; tail -3 gpurun_out/r2_kernel_sweep.log
python tools/loop_bench.py 125000 100 auto,stream > gpurun_out/r2_closed_loop_125k_100.log 2>&1; cat gpurun_out/r2_closed_loop_125k_100.log
