# usage: bash tools/mgN.sh <N>   -- bench lines of cfg3 (weak, strong), cfg4, cfg5 on N GPUs of one box into gpurun_out/r02s/
N=$1
set -x
mkdir -p gpurun_out/r02s
if [ "$N" = "1" ]; then TR="python"; else TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"; fi
timeout 400 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r02s/cfg3_weak_n$N.json 2> gpurun_out/r02s/cfg3_weak_n$N.err
timeout 400 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --no-parity --scaling strong > gpurun_out/r02s/cfg3_strong_n$N.json 2> gpurun_out/r02s/cfg3_strong_n$N.err
timeout 400 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --workload cfg4 > gpurun_out/r02s/cfg4_n$N.json 2> gpurun_out/r02s/cfg4_n$N.err
timeout 400 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --workload cfg5 > gpurun_out/r02s/cfg5_n$N.json 2> gpurun_out/r02s/cfg5_n$N.err
python - <<P
import json
for f in ('cfg3_weak','cfg3_strong','cfg4','cfg5'):
    try:
        for line in open('gpurun_out/r02s/%s_n$N.json' % f):
            if line.startswith('{'):
                d = json.loads(line); print(f, $N, round(d['value']), round(d['ms_per_step'], 3), round(d['e2e']['value']), d['e2e'].get('ms_per_step'))
    except Exception as e:
        print(f, 'ERR', e)
P
