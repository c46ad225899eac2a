python -m pytest tests -m gpu -x -q -s > gpurun_out/r2_gputest3.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_gputest3.log; grep "contact search\]" gpurun_out/r2_gputest3.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_r2c.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r2c.json')); b=d['configs3_bekker_grid']; print(b['value'], b['kernel_ms'], b['with_contact_search'])"
