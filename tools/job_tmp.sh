set -x
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_overlap.py -x -q -k "detect or fused" 2>&1 | tail -3
bash tools/gpu_job.sh r3d quick | tail -3 | cut -c1-300
