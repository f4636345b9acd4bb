#!/bin/bash
# round 2, call F: full suite after packed3 / DXT decode tables / decode scratch / wire tests; bench; JPEG decode host breakdown; converter sweep
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_f.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_f.log
tail -6 gpurun_out/pytest_f.log | cut -c1-600
timeout 900 python bench.py > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_f.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_f.json').read())
print("dxt1", d['value'], d['roofline']['us_per_launch'], "e2e", d['e2e']['value'])
for k,v in d['workloads'].items():
    if 'error' in v: print(k, v); continue
    print(k, "%.0f fps"%v['value'], "%.1f us"%(v['ms_per_frame']*1e3), "e2e %.0f"%v['e2e']['value'])
ex=d.get('extra',{})
for sec in ('kernels','decode'):
    for k,v in ex.get(sec,{}).items(): print(sec,k,{a:(round(b,2) if isinstance(b,float) else b) for a,b in v.items()} if isinstance(v,dict) else v)
PY
UGB200_JPEG_TIMING=1 timeout 300 python tools/profile_target.py jpegdec > gpurun_out/jpegdec_timing.txt 2>&1; tail -30 gpurun_out/jpegdec_timing.txt
timeout 600 python tools/pixfmt_sweep.py > gpurun_out/pixfmt_sweep.txt 2>&1; tail -70 gpurun_out/pixfmt_sweep.txt | cut -c1-200
