#!/bin/bash
# round 2, call D: full parity suite (incl. real-ABI modules through the reference framework), bench, ncu: full capture of the fused JPEG kernels + in-stream traffic
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_d.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_d.log
tail -6 gpurun_out/pytest_d.log | cut -c1-600
timeout 900 python bench.py > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_d.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_d.json').read())
print("dxt1", d['value'], d['roofline']['us_per_launch'], "e2e", d['e2e']['value'], "cpu", d.get('cpu_baseline',{}).get('value'))
for k,v in d['workloads'].items():
    if 'error' in v: print(k, v); continue
    r=v['roofline']; print(k, "%.0f fps"%v['value'], "%.1f us"%(v['ms_per_frame']*1e3), "frac %.3f"%r['frac'], "e2e %.0f"%v['e2e']['value'], "cpu", v.get('cpu_baseline',{}).get('value'), {x:round(r[x],1) for x in r if x.startswith('us_')})
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:jpeg_fused -s 4 -c 1 -o gpurun_out/prof_jpeg_uyvy -f python tools/profile_target.py jpeg > gpurun_out/ncu_jpeg_uyvy.log 2>&1; tail -2 gpurun_out/ncu_jpeg_uyvy.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:jpeg_fused -s 4 -c 1 -o gpurun_out/prof_jpeg_rgb -f python tools/profile_target.py jpeg_rgb > gpurun_out/ncu_jpeg_rgb.log 2>&1; tail -2 gpurun_out/ncu_jpeg_rgb.log
for k in dxt1 dxt6 p010 jpeg jpeg_rgb; do
  timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none --csv --log-file gpurun_out/traffic_$k.csv python tools/profile_target.py $k stream > gpurun_out/traffic_$k.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
