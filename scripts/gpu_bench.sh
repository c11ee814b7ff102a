#!/bin/bash
mkdir -p gpurun_out
N=${1:-1}
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
fi
echo "exit $?"; tail -5 gpurun_out/bench_${N}gpu.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_${N}gpu.json').read().strip().splitlines()[-1])
    print(json.dumps({k:d.get(k) for k in ('value','ms_per_step','e2e','gpu_launches','clocks','stages')}))
    print(json.dumps(d.get('roofline')))
    for k,v in (d.get('configs') or {}).items(): print(k, json.dumps(v)[:700])
    print(json.dumps(d.get('cpu_baseline'))); print(json.dumps(d.get('ref_cuda_build')))
except Exception as e: print('parse failed', e)
PY
