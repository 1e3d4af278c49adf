set -x
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_overlap.py -x -q -k "stem or branches or fused" 2>&1 | tail -3
bash tools/gpu_job.sh r4c quick | tail -3 | cut -c1-300
python bench.py --config 0 --no-cpu-baseline | cut -c1-200
