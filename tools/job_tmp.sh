set -x
python -m pytest tests/test_gpu_kernels.py -x -q -k "stem" 2>&1 | tail -3
bash tools/gpu_job.sh r4b quick | tail -3 | cut -c1-300
