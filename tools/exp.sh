python -m pytest tests/test_gpu_conv_tc.py -x -q 2>&1 | tail -4
python tools/conv_bench.py --reps 6 --only "Det.cv2" --prof 2>&1 | grep -v "^$" | head -12
echo "== streaming"; python tools/conv_bench.py --reps 6 --only "Det.cv2"
echo "== no streaming"; python tools/conv_bench.py --reps 6 --only "Det.cv2" --nostream
python bench.py --no-cpu-baseline --no-latency --steps 30 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['ms_per_step'], d['roofline']['classes']['fce_conv2d'])"
