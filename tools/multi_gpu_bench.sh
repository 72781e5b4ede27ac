# multi-GPU bench lines of the configs north_star names (no pytest): usage multi_gpu_bench.sh N "1 3 4"
N=${1:-2}
for c in ${2:-1 3 4}; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --config $c $( [ $c = 4 ] && echo "--steps 2" ) 2> gpurun_out/r2_bench_c${c}_${N}gpu.err | grep '^{' > gpurun_out/r2_bench_c${c}_${N}gpu.json
tail -c 300 gpurun_out/r2_bench_c${c}_${N}gpu.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_c${c}_${N}gpu.json').read().strip().splitlines()[-1])
print('config $c N=$N', round(d['value']), 'e2e', round(d['e2e']['value']), d['scaling'], d['config']['batch_per_gpu'], d['kernel'], 'ms', round(d['ms_per_step'],2))
"
done
