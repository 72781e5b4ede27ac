N=${1:-2}
python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/nccl_gather_check.py 4099 2>&1 | grep nccl_gather_check > gpurun_out/r2_nccl_gather_${N}gpu.json; cat gpurun_out/r2_nccl_gather_${N}gpu.json
for c in 1 3 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --config $c $( [ $c = 4 ] && echo "--steps 2" ) 2> gpurun_out/r2_bench_c${c}_${N}gpu.err | grep '^{' > gpurun_out/r2_bench_c${c}_${N}gpu.json
tail -c 300 gpurun_out/r2_bench_c${c}_${N}gpu.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_c${c}_${N}gpu.json').read().strip().splitlines()[-1])
print('config $c N=$N', round(d['value']), 'e2e', round(d['e2e']['value']), d['scaling'], d['config']['batch_per_gpu'], d['kernel'], 'ms', round(d['ms_per_step'],2), d.get('all_rank_stats'))
"
done
