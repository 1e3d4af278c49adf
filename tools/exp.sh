python -m pytest tests/test_gpu_overlap.py tests/test_gpu_preprocess.py -x -q 2>&1 | tail -4
python bench.py --no-cpu-baseline --no-latency --steps 30 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('overlap ', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])"
python bench.py --no-cpu-baseline --no-latency --steps 30 --no-overlap | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('inline  ', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])"
