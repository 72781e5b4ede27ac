# where does the stream mapping (with its iteration budget + hand-over, i.e. as the engine would pick it) overtake the shared-memory mappings?
for b in 24576 32768 49152 65536 98304 131072; do
  python tools/prof_solve.py warp LBMPC 50 $b 3 2>&1 | grep kernel_ms
  LBMPC_STREAM_MIN_BATCH=1 python tools/prof_solve.py auto LBMPC 50 $b 3 2>&1 | grep kernel_ms
done
for b in 8192 16384 24576 32768 49152; do
  python tools/prof_solve.py cta LBMPC 200 $b 3 2>&1 | grep kernel_ms
  LBMPC_STREAM_MIN_BATCH_LONG=1 python tools/prof_solve.py auto LBMPC 200 $b 3 2>&1 | grep kernel_ms
done
for b in 65536 131072 262144; do
  python tools/prof_solve.py cta LMPC 50 $b 3 2>&1 | grep kernel_ms
  python tools/prof_solve.py stream LMPC 50 $b 3 2>&1 | grep kernel_ms
done
