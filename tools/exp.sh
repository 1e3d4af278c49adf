python -m pytest tests/test_gpu_kernels.py -x -q -k "stem" 2>&1 | tail -3
python bench.py --no-cpu-baseline --no-latency --steps 30 --kernel-times gpurun_out/kt.csv | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['ms_per_step'], d['roofline']['classes']['fce_stem_conv'])"
