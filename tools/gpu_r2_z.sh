#!/bin/bash
# round 2, call Z: N = 2 bench line of the final state, launched as the driver launches it (torchrun), + the multi-device module tests
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi_L.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_j_2gpu.json 2> gpurun_out/bench_j_2gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_j_2gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_j_2gpu.json').read())
print("n_gpus", d['n_gpus'], "dxt1", d['value'], d['roofline']['us_per_launch'], "e2e", d['e2e']['value'], d['e2e'].get('h2d_GBps_per_gpu'))
for k,v in d['workloads'].items():
    if 'error' in v: print(k, v); continue
    print(k, "%.0f fps"%v['value'], "e2e %.0f"%v['e2e']['value'])
PY
