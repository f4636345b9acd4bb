#!/bin/bash
# round 2, call L: full suite with the real-ABI decompress modules, the kernels.cu entry points and the staged line kernel; converter sweep direct vs staged; bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_l.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_l.log
tail -6 gpurun_out/pytest_l.log | cut -c1-600
timeout 900 python tools/pixfmt_sweep.py 0,1 > gpurun_out/pixfmt_sweep_l.txt 2>&1; tail -70 gpurun_out/pixfmt_sweep_l.txt | cut -c1-200
timeout 900 python bench.py > gpurun_out/bench_l.json 2> gpurun_out/bench_l.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_l.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_l.json').read())
print("dxt1", d['value'], d['roofline']['us_per_launch'], "e2e", d['e2e']['value'])
for k,v in d['workloads'].items():
    if 'error' in v: print(k, v); continue
    print(k, "%.0f fps"%v['value'], "%.1f us"%(v['ms_per_frame']*1e3), "e2e %.0f"%v['e2e']['value'])
PY
