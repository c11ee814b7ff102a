#!/bin/bash
# full GPU check: every gpu test, smoke, a bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1
echo "pytest gpu exit $?"; grep -E "passed|failed|FAILED|mismatch|C2 full|statistics|fast vs exact" gpurun_out/pytest_gpu.log | cut -c1-300 | head -30
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','stages','roofline') if k in d})
PY
