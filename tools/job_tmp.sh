python tools/conv_bench.py --prof > gpurun_out/cb_prof_r2h.log 2>&1
python tools/conv_bench.py --dbg 1 > gpurun_out/cb_dbg1_r2h.log 2>&1
python tools/conv_bench.py --dbg 6 > gpurun_out/cb_dbg6_r2h.log 2>&1
cat gpurun_out/cb_prof_r2h.log
