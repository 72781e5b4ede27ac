set -x
python bench.py > gpurun_out/r2_bench_c1.json 2> gpurun_out/r2_bench_c1.err; tail -c 600 gpurun_out/r2_bench_c1.err
python bench.py --config 2 > gpurun_out/r2_bench_c2.json 2> gpurun_out/r2_bench_c2.err; tail -c 600 gpurun_out/r2_bench_c2.err
python bench.py --config 3 > gpurun_out/r2_bench_c3.json 2> gpurun_out/r2_bench_c3.err; tail -c 600 gpurun_out/r2_bench_c3.err
python bench.py --config 4 --steps 2 > gpurun_out/r2_bench_c4.json 2> gpurun_out/r2_bench_c4.err; tail -c 600 gpurun_out/r2_bench_c4.err
for f in gpurun_out/r2_bench_c?.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1])
print('$f', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'kernel', d['kernel'], 'ms', d['ms_per_step'], 'cpu', d['cpu_baseline']['value'], d.get('latency_us_single_qp'), d.get('e2e_pageable',{}).get('value'), d['status_counts'])
"; done
