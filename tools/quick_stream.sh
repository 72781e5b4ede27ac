for b in 131072 262144; do LBMPC_STREAM_WARPS=10 python tools/prof_solve.py stream LBMPC 50 $b 3; python tools/prof_solve.py stream LBMPC 50 $b 3; done
