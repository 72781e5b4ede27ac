# stream mapping: timings at the sizes that matter + the GPU tests that force it
for cfg in "LBMPC 50 65536" "LBMPC 50 262144" "LMPC 50 262144" "LBMPC 200 65536"; do python tools/prof_solve.py stream $cfg 3; done 2>&1 | grep kernel_ms | tee gpurun_out/stream_check.log
python tools/prof_solve.py auto LBMPC 200 65536 3 2>&1 | grep kernel_ms | tee -a gpurun_out/stream_check.log
python tools/prof_solve.py mixed LBMPC 50 262144 3 2>&1 | grep kernel_ms | tee -a gpurun_out/stream_check.log
python -m pytest tests -m gpu -q -x -k "stream or kernel or budget or sqp or closed_loop" 2>&1 | tail -3 | tee -a gpurun_out/stream_check.log
