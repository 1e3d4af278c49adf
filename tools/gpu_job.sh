#!/bin/bash
# Round job on the B200 box: tests, bench (both arms), ncu launch list of the bench command, and ncu --set full of
# one eager step for a SELECTION of kernels, exported to CSV on the box (gpurun_out is capped at 64 MiB).
# usage: bash tools/gpu_job.sh <tag> [quick|notest]     quick = tests + bench only
TAG=${1:-r1x}
set -x
if [ "$2" != "notest" ]; then
  python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"
  tail -3 gpurun_out/pytest_$TAG.log
fi
python bench.py --kernel-times gpurun_out/ktimes_$TAG.csv > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
if [ "$2" == "quick" ]; then cat gpurun_out/bench_$TAG.json; exit 0; fi
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"
# the "kernel to beat" (SURVEY 8d): the same graph through stock PyTorch (cuDNN, bf16 channels_last) on this GPU
timeout 300 python tests/torch_gpu_baseline.py > gpurun_out/torch_eager_$TAG.json 2> gpurun_out/torch_eager_$TAG.err; echo "torch eager rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/ncu_l_$TAG.log 2>&1; echo "ncu launches rc=$?"
# full-set capture: first 14 conv launches (stem .. layer 4: halo, im2col s2 and 1x1 kernels) + every non-conv kernel
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:'conv_tc|conv_halo' -c 14 -f -o /tmp/conv_$TAG python tools/profile_step.py --nodes gpurun_out/nodes_$TAG.csv \
  > gpurun_out/ncu_conv_$TAG.log 2>&1; echo "ncu conv rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:'dwconv|psa|nms|coord_pool|coordatt_mlp|gate|bifpn|stem|decode|sppf|upsample|strip_attn|simt' -c 40 -f -o /tmp/bw_$TAG \
  python tools/profile_step.py > gpurun_out/ncu_bw_$TAG.log 2>&1; echo "ncu bw rc=$?"
# DRAM bytes of every launch of one eager step -> profiles/traffic.json (bench.py's roofline.traffic)
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/traffic_$TAG.csv python tools/profile_step.py > gpurun_out/ncu_t_$TAG.log 2>&1; echo "ncu traffic rc=$?"
python tools/ncu_traffic.py gpurun_out/traffic_$TAG.csv gpurun_out/traffic_$TAG.json "$(python -c 'import bench; print(bench.WORKLOAD_NAME)')" > /dev/null
for n in conv bw; do
  ncu -i /tmp/${n}_$TAG.ncu-rep --page raw --csv > gpurun_out/ncu_${n}_${TAG}_raw.csv 2>/dev/null
  ls -la /tmp/${n}_$TAG.ncu-rep
done
# keep the conv report itself when it is small enough to travel
sz=$(stat -c %s /tmp/conv_$TAG.ncu-rep 2>/dev/null || echo 0)
if [ "$sz" -gt 0 ] && [ "$sz" -lt 30000000 ]; then cp /tmp/conv_$TAG.ncu-rep gpurun_out/; fi
sz=$(stat -c %s /tmp/bw_$TAG.ncu-rep 2>/dev/null || echo 0)
if [ "$sz" -gt 0 ] && [ "$sz" -lt 20000000 ]; then cp /tmp/bw_$TAG.ncu-rep gpurun_out/; fi
du -sh gpurun_out
