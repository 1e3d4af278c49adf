set -x
python bench.py --kernel-times gpurun_out/ktimes_r5.csv > gpurun_out/bench_r5.json 2> gpurun_out/bench_r5.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r5.json 2> gpurun_out/bench_ref_r5.err; echo "ref rc=$?"
python bench.py --config 0 --no-cpu-baseline > gpurun_out/bench_cfg0_r5.json 2>/dev/null; echo "cfg0 rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
cut -c1-250 gpurun_out/bench_r5.json
