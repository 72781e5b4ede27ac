# small stream launches spread over all SMs (2 / 4 warps per CTA) vs the shared-memory mappings
for b in 2048 4096 8192 12288 16384 20480; do
  python tools/prof_solve.py cta LBMPC 200 $b 3 2>&1 | grep kernel_ms
  LBMPC_STREAM_MIN_BATCH_LONG=1 python tools/prof_solve.py auto LBMPC 200 $b 3 2>&1 | grep kernel_ms
  LBMPC_STREAM_SPREAD=0 LBMPC_STREAM_MIN_BATCH_LONG=1 python tools/prof_solve.py auto LBMPC 200 $b 3 2>&1 | grep kernel_ms
done
for b in 4096 8192 16384 24576; do
  python tools/prof_solve.py warp LBMPC 50 $b 3 2>&1 | grep kernel_ms
  LBMPC_STREAM_MIN_BATCH=1 python tools/prof_solve.py auto LBMPC 50 $b 3 2>&1 | grep kernel_ms
done
