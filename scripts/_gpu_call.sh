python bench.py > gpurun_out/bench_r2e.json 2> gpurun_out/bench_r2e.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_r2e.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_r2.log 2>&1; echo "ncu launches rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2e_ref.json 2> gpurun_out/bench_r2e_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_r2e_ref.json | cut -c1-400
