#!/bin/bash
# round 2, call B: JPEG changes (scan kernel, word-wise stuffing, staged RGB tile) + DXT1 skew variants
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_jpeg.py tests/test_vcompress.py tests/test_jpeg_decode.py tests/test_jpeg_wire.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_b.log
tail -8 gpurun_out/pytest_b.log | cut -c1-600
UGB200_DXT1_SKEW=1 timeout 600 python -m pytest tests/test_dxt_gpu.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_b_skew.log 2>&1; echo "pytest skew rc=$?" >> gpurun_out/pytest_b_skew.log
tail -4 gpurun_out/pytest_b_skew.log | cut -c1-400
timeout 300 tools/exp_dxt > gpurun_out/exp_dxt_b.txt 2>&1; grep -E "^d1" gpurun_out/exp_dxt_b.txt | sort -k12 -n | head -30
timeout 300 tools/exp_dxt_o1 > gpurun_out/exp_dxt_b_o1.txt 2>&1; grep -E "^d1" gpurun_out/exp_dxt_b_o1.txt | sort -k12 -n | head -12
timeout 900 python bench.py > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_b.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_b.json').read())
print("dxt1", d['value'], d['roofline']['us_per_launch'], "e2e", d['e2e']['value'], "cpu", d.get('cpu_baseline',{}).get('value'), d.get('cpu_baseline',{}).get('cores'))
for k,v in d['workloads'].items():
    if 'error' in v: print(k, v); continue
    r=v['roofline']; print(k, "%.0f fps"%v['value'], "%.1f us"%(v['ms_per_frame']*1e3), "frac %.3f"%r['frac'], "e2e %.0f"%v['e2e']['value'], "cpu", v.get('cpu_baseline',{}).get('value'), {x:round(r[x],1) for x in r if x.startswith('us_')})
PY
