python bench.py > gpurun_out/bench_r2h.json 2> gpurun_out/bench_r2h.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2h_ref.json 2> gpurun_out/bench_r2h_ref.err; echo "ref rc=$?"; tail -c 600 gpurun_out/bench_r2h_ref.json
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python scripts/quick_gpu.py 8192 600 200 2 2>&1 | grep "grad=" 
python scripts/quick_gpu.py 16384 600 200 2 2>&1 | grep "grad="
