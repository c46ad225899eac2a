python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python scripts/ab_variants.py base main
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
