set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r1j.log 2>&1; echo "pytest rc=$?"
python bench.py --kernel-times gpurun_out/ktimes_r1j.csv > gpurun_out/bench_r1j.json 2> gpurun_out/bench_r1j.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r1j.json 2> gpurun_out/bench_ref_r1j.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r1j.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l_r1j.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/step_r1j python tools/profile_step.py --nodes gpurun_out/nodes_r1j.csv > gpurun_out/ncu_full_r1j.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/
