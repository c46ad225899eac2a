python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python scripts/ab_variants.py base main
ncu --set full --clock-control none --import-source on -k regex:rollout_bwd -c 1 -o gpurun_out/prof_bwd_v4 -f python scripts/quick_gpu.py 4096 20 200 1 > gpurun_out/ncu_bwd_v4.log 2>&1; echo "ncu rc=$?"
