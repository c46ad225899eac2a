python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python scripts/ab_variants.py base main
python scripts/ncu_targets.py once > gpurun_out/ncu_targets_plain.log 2>&1; echo "targets rc=$?"
ncu --set full --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,sm__inst_executed_pipe_fp64.sum --clock-control none --import-source on -k regex:rollout -c 4 -f -o gpurun_out/prof_r2g python scripts/ncu_targets.py once > gpurun_out/ncu_r2g.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_r2g.log
