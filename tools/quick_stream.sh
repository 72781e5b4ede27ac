for e in 14 18 22 26 32; do LBMPC_STREAM_EVICT=$e python tools/prof_solve.py stream LBMPC 200 65536 3; done
for e in 14 18 24; do LBMPC_STREAM_EVICT=$e python tools/prof_solve.py stream LBMPC 50 65536 3; done
for e in 10 14 18 24; do LBMPC_STREAM_EVICT=$e python tools/prof_solve.py stream LBMPC 50 262144 3; done
for e in 12 18; do LBMPC_STREAM_EVICT=$e python tools/prof_solve.py stream LBMPC 50 131072 3; done
python tools/prof_solve.py warp LBMPC 50 131072 3
