set -x
python -m pytest tests/test_gpu_kernels.py -x -q -k "weighted" 2>&1 | tail -3
bash tools/gpu_job.sh r3e quick | tail -3 | cut -c1-300
