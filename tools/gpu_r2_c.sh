#!/bin/bash
# round 2, call C: JPEG with TMA bulk staging + PDL; 8-CTA experiment; bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_jpeg.py tests/test_vcompress.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_c.log
tail -6 gpurun_out/pytest_c.log | cut -c1-600
timeout 600 python bench.py --only uyvy_jpeg_8k_q90,rgb_jpeg_8k_q90 > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_c.err
UGB200_JPEG_EIGHT=1 timeout 600 python bench.py --only uyvy_jpeg_8k_q90,rgb_jpeg_8k_q90 > gpurun_out/bench_c8.json 2> gpurun_out/bench_c8.err; echo "bench8 rc=$?"; tail -3 gpurun_out/bench_c8.err
python - <<'PY'
import json
for f in ('bench_c','bench_c8'):
    d=json.loads(open(f'gpurun_out/{f}.json').read())
    print(f, "dxt1", d['value'], d['roofline']['us_per_launch'])
    for k,v in d['workloads'].items():
        if 'error' in v: print(k, v); continue
        r=v['roofline']; print(k, "%.0f fps"%v['value'], "single %.1f us two-stream %.1f us"%(v['single_stream_ms_per_frame']*1e3, v['two_stream_ms_per_frame']*1e3), "e2e %.0f"%v['e2e']['value'], {x:round(r[x],1) for x in r if x.startswith('us_')})
PY
UGB200_JPEG_EIGHT=1 timeout 600 python -m pytest tests/test_jpeg.py -m gpu -q -x --timeout 600 -k "equals_oracle or serial_route" > gpurun_out/pytest_c8.log 2>&1; tail -3 gpurun_out/pytest_c8.log
