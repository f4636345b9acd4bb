#!/bin/bash
# round 2, call Y (final state of the round, same evidence as call T): all GPU parity tests, smoke, bench line with every extra, reference arm, launch list of the bench command, fused JPEG kernel capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_j.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_j.log
tail -4 gpurun_out/pytest_j.log | cut -c1-300
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_j.log 2>&1; tail -1 gpurun_out/smoke_j.log
timeout 900 python bench.py > gpurun_out/bench_j.json 2> gpurun_out/bench_j.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_j.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_j.json').read())
print("dxt1", d['value'], d['roofline']['us_per_launch'], "e2e", d['e2e']['value'])
for k,v in d['workloads'].items():
    if 'error' in v: print(k, v); continue
    print(k, "%.0f fps"%v['value'], "%.1f us"%(v['ms_per_frame']*1e3), "e2e %.0f"%v['e2e']['value'])
print(json.dumps(d['extra'].get('decode'))[:900])
PY
timeout 600 python bench.py --impl reference > gpurun_out/bench_j_reference.json 2> gpurun_out/bench_j_reference.err; echo "reference arm rc=$?"; cut -c1-400 gpurun_out/bench_j_reference.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_j.csv python bench.py --steps 2 --warmup 3 --no-extra > gpurun_out/bench_under_ncu.log 2>&1; tail -2 gpurun_out/launches_j.csv | cut -c1-200
