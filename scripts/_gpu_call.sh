python scripts/quick_bekker.py 8192 100 | tail -2
AVS_LIB=$PWD/auvsl_neural_ode_b200/libavsim_bc.so python scripts/quick_bekker.py 8192 100 | tail -2
AVS_LIB=$PWD/auvsl_neural_ode_b200/libavsim_bc.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bekker or grid or kats or contact" 2>&1 | tail -3
