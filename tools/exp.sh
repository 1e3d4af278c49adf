python bench.py --no-cpu-baseline --no-latency --steps 30 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench cfg1', d['value'], d['ms_per_step'])"
for c in 0 2 3 4; do
  python bench.py --config $c --no-cpu-baseline --no-latency --steps 10 --warmup 3 > gpurun_out/bench_cfg$c.json 2> gpurun_out/bench_cfg$c.err
  python -c "
import json,sys
try:
    d=json.loads(open('gpurun_out/bench_cfg$c.json').read()); r=d['roofline']
    print('cfg$c', d['value'], 'img/s', d['ms_per_step'], 'ms/step  e2e', d['e2e']['value'], ' arena MB', d['config']['arena_mb'], ' conv', r['classes']['fce_conv2d'])
except Exception as e:
    print('cfg$c FAILED', e); print(open('gpurun_out/bench_cfg$c.err').read()[-1500:])
"
done
