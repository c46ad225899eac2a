timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py > gpurun_out/bench_r2g.json 2> gpurun_out/bench_r2g.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2g.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','clocks')})
print(d['roofline'])
for k in ('configs1_forward_only','configs3_bekker_grid','strong_scaling'):
    print(k, {kk:vv for kk,vv in d[k].items() if kk in ('value','kernel_ms','ms_per_step')})
PY
