python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stream or kernels" 2>&1 | tail -3
for b in 65536 262144; do python tools/prof_solve.py stream LBMPC 50 $b 4; python tools/prof_solve.py mixed LBMPC 50 $b 4; done
