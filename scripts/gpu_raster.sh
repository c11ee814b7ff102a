#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -m gpu -q -k "render or refiner or full_size or smoke" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout -k 10 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -s 18 -c 13 --csv --log-file gpurun_out/launches2.csv python scripts/profile_step.py 3 > gpurun_out/launches2.log 2>&1
timeout -k 10 300 python scripts/time_stages.py 512 6 > gpurun_out/time_persistent.json 2> gpurun_out/time_persistent.err
tail -3 gpurun_out/pytest_gpu.log; python - <<'PY'
import csv
lines=[l for l in open('gpurun_out/launches2.csv') if l.startswith('"')]
for row in csv.DictReader(lines):
    print(row['ID'], row['Kernel Name'][:40], row['Metric Name'], row['Metric Value'], row['Metric Unit'])
PY
cut -c1-330 gpurun_out/time_persistent.json
