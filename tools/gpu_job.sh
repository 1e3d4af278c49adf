#!/bin/bash
# Profiling job on the B200 box (after the bench command has run clean): ncu launch list of the bench command, ncu
# --set full of ONE eager step for a selection of kernels (exported to CSV on the box: gpurun_out is capped at 64 MiB)
# and ncu DRAM bytes of every launch of one eager step -> profiles/traffic.json.
# usage: bash tools/gpu_job.sh <tag> [batch]      (batch: override the workload's batch for the --set full captures)
TAG=${1:-r02x}
PB=${2:-64}
set -x
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-configs --no-latency"
$BENCH > gpurun_out/bench_short_$TAG.json 2> gpurun_out/bench_short_$TAG.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_$TAG.csv \
  $BENCH > gpurun_out/ncu_l_$TAG.log 2>&1; echo "ncu launches rc=$?"
# full-set capture: the first conv launches of the step (stem .. layer 6: pair 3x3/s2, 1x1, strip kernels) + every non-conv kernel
python tools/profile_step.py --batch $PB --nodes gpurun_out/nodes_$TAG.csv > gpurun_out/profile_step_$TAG.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:'conv_tc|conv_halo|conv_dwpw|conv_stem2|conv_chain' -c 34 -f -o /tmp/conv_$TAG python tools/profile_step.py --batch $PB \
  > gpurun_out/ncu_conv_$TAG.log 2>&1; echo "ncu conv rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:'dwconv|psa|nms|coord_pool|coordatt_mlp|gate|bifpn|stem|decode|sppf|upsample|strip_attn|simt' -c 40 -f -o /tmp/bw_$TAG \
  python tools/profile_step.py --batch $PB > gpurun_out/ncu_bw_$TAG.log 2>&1; echo "ncu bw rc=$?"
# DRAM bytes of every launch of one eager step at the FULL batch -> traffic json (bench.py's roofline.traffic)
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/traffic_$TAG.csv python tools/profile_step.py > gpurun_out/ncu_t_$TAG.log 2>&1; echo "ncu traffic rc=$?"
python tools/ncu_traffic.py gpurun_out/traffic_$TAG.csv gpurun_out/traffic_$TAG.json "$(python -c 'import bench; print(bench.WORKLOAD_NAME)')" > /dev/null
for n in conv bw; do
  ncu -i /tmp/${n}_$TAG.ncu-rep --page raw --csv > gpurun_out/ncu_${n}_${TAG}_raw.csv 2>/dev/null
  ls -la /tmp/${n}_$TAG.ncu-rep
done
sz=$(stat -c %s /tmp/conv_$TAG.ncu-rep 2>/dev/null || echo 0)
if [ "$sz" -gt 0 ] && [ "$sz" -lt 30000000 ]; then cp /tmp/conv_$TAG.ncu-rep gpurun_out/; fi
du -sh gpurun_out
