python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python scripts/ncu_targets.py once > gpurun_out/ncu_targets_plain.log 2>&1; echo "targets rc=$?"
ncu --set full --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,sm__inst_executed_pipe_fp64.sum --clock-control none --import-source on -k regex:rollout -c 4 -f -o gpurun_out/prof_r2i python scripts/ncu_targets.py once > gpurun_out/ncu_r2i.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_r2i.log
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; echo "bench plain rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2i.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_r2i.log 2>&1; echo "launch list rc=$?"
