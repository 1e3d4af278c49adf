timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'stem_fused|dwconv_tma|coord_pool_kernel|decode_kernel' -c 6 -f -o /tmp/k python tools/profile_step.py > gpurun_out/ncu_k.log 2>&1
ncu -i /tmp/k.ncu-rep --page raw --csv > gpurun_out/k_raw.csv 2>/dev/null
ncu -i /tmp/k.ncu-rep --page source --csv --print-source sass > gpurun_out/k_src_sass.csv 2>/dev/null
ls -la gpurun_out/k_*
