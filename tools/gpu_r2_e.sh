#!/bin/bash
# round 2, call E: pipelined entropy loop + L2 prefetch; lavc bridge tests; bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_jpeg.py tests/test_lavc.py tests/test_vcompress.py tests/test_real_module.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_e.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_e.log
tail -8 gpurun_out/pytest_e.log | cut -c1-700
timeout 600 python bench.py --only uyvy_jpeg_8k_q90,rgb_jpeg_8k_q90 > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_e.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_e.json').read())
for k,v in d['workloads'].items():
    if 'error' in v: print(k, v); continue
    r=v['roofline']; print(k, "%.0f fps"%v['value'], "single %.1f us two-stream %.1f us"%(v['single_stream_ms_per_frame']*1e3, v['two_stream_ms_per_frame']*1e3), "e2e %.0f"%v['e2e']['value'], {x:round(r[x],1) for x in r if x.startswith('us_')})
PY
