# usage: bash tools/mgN2.sh <N> [strong]  -- bench lines of cfg3 (weak [, strong]), cfg4 / cfg5 (engines tc and tc2) on N GPUs of one box into gpurun_out/r03s/
N=$1
mkdir -p gpurun_out/r03s
if [ "$N" = "1" ]; then TR="python"; else TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"; fi
[ "$N" = "8" ] && nvidia-smi topo -m > gpurun_out/r03s/topo_n8.txt 2>&1
timeout 400 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r03s/cfg3_weak_n$N.json 2> gpurun_out/r03s/cfg3_weak_n$N.err
if [ "$2" = "strong" ]; then
timeout 400 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --no-parity --scaling strong > gpurun_out/r03s/cfg3_strong_n$N.json 2> gpurun_out/r03s/cfg3_strong_n$N.err
fi
for e in tc tc2; do for w in cfg4 cfg5; do
timeout 400 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --workload $w --f32-engine $e > gpurun_out/r03s/${w}_${e}_n$N.json 2> gpurun_out/r03s/${w}_${e}_n$N.err
done; done
python - <<P
import json, glob
for f in sorted(glob.glob('gpurun_out/r03s/*_n$N.json')):
    try:
        for line in open(f):
            if line.startswith('{'):
                d = json.loads(line); print(f.split('/')[-1], round(d['value']), round(d['ms_per_step'], 3), round(d['e2e']['value']), d['e2e'].get('ms_per_step'), d['clocks']['sm_mhz'], d['clocks']['reasons'])
    except Exception as e:
        print(f, 'ERR', e)
P
