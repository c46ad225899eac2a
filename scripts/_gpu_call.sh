python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python scripts/ab_variants.py main eg0
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
