set -x
python -m pytest tests/test_gpu_overlap.py -x -q 2>&1 | tail -3
bash tools/gpu_job.sh r4d quick | tail -3 | cut -c1-200
