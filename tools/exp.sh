python -m pytest tests/test_gpu_parity.py -x -q -k "nms" 2>&1 | tail -2
python tools/nms_bench.py 2>&1 | tail -3
python bench.py --no-cpu-baseline --no-latency --steps 30 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['ms_per_step'], d['roofline']['classes']['fce_nms'])"
