python bench.py > gpurun_out/bench_r2f.json 2> gpurun_out/bench_r2f.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_r2f.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_r2.log 2>&1; echo "ncu launches rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
