set -x
python -m pytest tests/test_gpu_kernels.py -x -q -k "detect" 2>&1 | tail -5
python -m pytest tests/test_gpu_overlap.py -x -q 2>&1 | tail -5
bash tools/gpu_job.sh r3c quick | tail -3 | cut -c1-300
