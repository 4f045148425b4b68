mkdir -p gpurun_out/r03f
for c in 8192 4096 2048; do
  timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-exact-mode --no-parity --chunk $c > gpurun_out/r03f/cfg3_c$c.json 2> gpurun_out/r03f/cfg3_c$c.err || tail -c 500 gpurun_out/r03f/cfg3_c$c.err
  python - gpurun_out/r03f/cfg3_c$c.json $c <<'P'
import json,sys
for line in open(sys.argv[1]):
    if line.startswith('{'):
        d=json.loads(line); e=d['e2e']
        print('chunk', sys.argv[2], 'value %.0f ms %.3f | e2e %.0f ms %.3f h2d %d d2h %d launches %s' % (d['value'], d['ms_per_step'], e['value'], e['ms_per_step'], e['h2d_bytes_per_step'], e['d2h_bytes_per_step'], e.get('gpu_launches')))
P
done
