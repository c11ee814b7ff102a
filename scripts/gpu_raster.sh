#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "render or division or refiner or c5 or full_size or smoke" -s 2>&1 | grep -E "passed|failed|FAILED|^E  " | cut -c1-300 | head
timeout 300 python scripts/time_step.py 512 10 | tee gpurun_out/time_step.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 24 --csv --log-file gpurun_out/launches_step.csv python scripts/time_step.py 512 1 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_step.csv')) if len(r)>5]
hdr=rows[0]; i=hdr.index('Kernel Name'); v=hdr.index('Metric Value')
for r in rows[1:25]: print(r[i][:50], r[v])
PY
