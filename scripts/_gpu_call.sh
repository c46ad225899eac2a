AVS_LIB=$PWD/auvsl_neural_ode_b200/libavsim_zf.so python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python scripts/ab_variants.py main zf
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
