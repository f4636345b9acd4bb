#!/bin/bash
# round 2, call T (state i evidence): all GPU parity tests, smoke, bench line with every extra, reference arm, launch list of the bench command, fused JPEG kernel capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_i.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_i.log
tail -4 gpurun_out/pytest_i.log | cut -c1-300
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_i.log 2>&1; tail -1 gpurun_out/smoke_i.log
timeout 900 python bench.py > gpurun_out/bench_i.json 2> gpurun_out/bench_i.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_i.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_i.json').read())
print("dxt1", d['value'], d['roofline']['us_per_launch'], "e2e", d['e2e']['value'])
for k,v in d['workloads'].items():
    if 'error' in v: print(k, v); continue
    print(k, "%.0f fps"%v['value'], "%.1f us"%(v['ms_per_frame']*1e3), "e2e %.0f"%v['e2e']['value'])
print(json.dumps(d['extra'].get('decode'))[:900])
PY
timeout 600 python bench.py --impl reference > gpurun_out/bench_i_reference.json 2> gpurun_out/bench_i_reference.err; echo "reference arm rc=$?"; cut -c1-400 gpurun_out/bench_i_reference.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_i.csv python bench.py --steps 2 --warmup 3 --no-extra > gpurun_out/bench_under_ncu.log 2>&1; tail -2 gpurun_out/launches_i.csv | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:jpeg_fused -s 4 -c 1 -o gpurun_out/prof_jpeg_uyvy_i -f python tools/profile_target.py jpeg > gpurun_out/ncu_jpeg_uyvy_i.log 2>&1; tail -1 gpurun_out/ncu_jpeg_uyvy_i.log
ls -la gpurun_out/*.ncu-rep
