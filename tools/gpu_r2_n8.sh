#!/bin/bash
# round 2: 8-GPU scaling check of the whole bench (product arm), launched as the driver does
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
cat /sys/fs/cgroup/cpu.max > gpurun_out/cpu8.txt 2>&1; nproc >> gpurun_out/cpu8.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "bench n8 rc=$?"; tail -5 gpurun_out/bench_n8.err | cut -c1-300
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_n8.json') if l.startswith('{')][-1])
    print("N", d['n_gpus'], "dxt1", d['value'], "us/launch", d['roofline']['us_per_launch'], "e2e", d['e2e']['value'], "h2d GB/s per gpu", d['e2e'].get('h2d_GBps_per_gpu'))
    for k,v in d['workloads'].items():
        if 'error' in v: print(k, v); continue
        print(k, "%.0f fps"%v['value'], "e2e %.0f"%v['e2e']['value'])
except Exception as e:
    print("parse failed", e)
PY
