python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "closed_loop or config5" 2>&1 | tail -8
python tools/loop_bench.py 125000 20 auto,stream
LBMPC_LOOP_CHUNK=5 python tools/loop_bench.py 125000 20 stream
