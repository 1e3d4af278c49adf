python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== bn cap 128"
FCE_BN_MAX=128 python tools/conv_bench.py --reps 6 --only "L4.cv2,L6,L7.cv2,L9,L10.cv2,L11.cv2"
FCE_BN_MAX=128 python bench.py --no-cpu-baseline --no-latency --steps 30 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench bn128', d['value'], d['ms_per_step'], d['roofline']['classes']['fce_conv2d'])"
python bench.py --no-cpu-baseline --no-latency --steps 30 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench base', d['value'], d['ms_per_step'], d['roofline']['classes']['fce_conv2d'])"
